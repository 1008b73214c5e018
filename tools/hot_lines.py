#!/usr/bin/env python3
"""Aggregate the ncu source page (`ncu -i X.ncu-rep --page source --csv`) by CUDA source line:
warp-stall samples (where the warps' time goes), executed instructions and the dominant stall reasons.

    ncu -i gpurun_out/<tag>_prof.ncu-rep --page source --csv > /tmp/src.csv
    python tools/hot_lines.py /tmp/src.csv [source.cu] [top]
"""
import collections
import csv
import sys


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


def main():
    path = sys.argv[1]
    src = sys.argv[2] if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(path)))
    # several kernels / launches may be concatenated: take the first table
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No')
    hdr = rows[start]
    col = {n: i for i, n in enumerate(hdr)}
    # the header has two "Source" columns (CUDA line text, SASS text): first is at index 1, SASS at 3
    stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
    by_line = collections.defaultdict(lambda: collections.Counter())
    text = {}
    total = collections.Counter()
    cur = None
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[0] == 'Line No':
            break
        if r[0] == '':                       # a SASS row of the current source line
            if cur is None:
                continue
            parts = r[3].split()
            op = parts[1] if parts and parts[0].startswith('@') and len(parts) > 1 else (parts[0] if parts else '')
            n = num(r[col['Instructions Executed']])
            if op[:4] in ('DFMA', 'DMUL', 'DADD'):
                by_line[cur]['fp64'] += n
                total['fp64'] += n
            continue
        try:
            cur = int(r[0])
        except ValueError:
            cur = None
            continue
        text[cur] = r[1]
        d = by_line[cur]
        d['samples'] += num(r[col['# Samples']])
        d['inst'] += num(r[col['Instructions Executed']])
        for n in stall_cols:
            v = num(r[col[n]])
            d[n] += v
            total[n] += v
        total['samples'] += num(r[col['# Samples']])
        total['inst'] += num(r[col['Instructions Executed']])
    print(f'total samples {total["samples"]}, warp instructions {total["inst"]}, of which FP64 arithmetic '
          f'{100 * total["fp64"] / max(1, total["inst"]):.1f}%')
    print('stall mix: ' + ', '.join(f'{n[6:]} {100 * total[n] / max(1, total["samples"]):.1f}%'
                                    for n in sorted(stall_cols, key=lambda n: -total[n])[:8]))
    print(f'{"line":>5} {"samples%":>8} {"inst%":>6} {"fp64%":>6}  top stalls / source')
    for line, d in sorted(by_line.items(), key=lambda kv: -kv[1]['samples'])[:top]:
        st = sorted(stall_cols, key=lambda n: -d[n])[:3]
        stalls = ' '.join(f'{n[6:]}={100 * d[n] / max(1, d["samples"]):.0f}%' for n in st)
        code = text.get(line, '').strip()[:90]
        print(f'{line:5d} {100 * d["samples"] / max(1, total["samples"]):8.2f} {100 * d["inst"] / max(1, total["inst"]):6.2f} '
              f'{100 * d["fp64"] / max(1, d["inst"]):6.0f}  {stalls} | {code}')


if __name__ == '__main__':
    main()
