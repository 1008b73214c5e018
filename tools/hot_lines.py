#!/usr/bin/env python3
"""Aggregate the ncu source page (`ncu -i X.ncu-rep --page source --csv`) by CUDA source line:
warp-stall samples (where the warps' time goes), executed instructions and the dominant stall reasons.

    ncu -i gpurun_out/<tag>_prof.ncu-rep --page source --csv > /tmp/src.csv
    python tools/hot_lines.py /tmp/src.csv [- | library.so] [top]

When the report holds no CUDA source (SASS-only page), pass the library the profile was taken with: the
SASS rows are then mapped to source lines through `nvdisasm -g` of its cubin (needs -lineinfo) and the line
text is read from the source file named there.
"""
import collections
import csv
import sys


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


def sass_only(rows, lib, top):
    """SASS-only source page: map instruction offsets to lines with nvdisasm."""
    import os
    import re
    import subprocess
    import tempfile
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    kernel = next(r[1] for r in rows[:start] if r and r[0] == 'Kernel Name')
    hdr = rows[start]
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
    data = []
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[0] == 'Address':
            break
        data.append(r)
    base = min(int(r[0], 16) for r in data)
    # nvdisasm of the same function: offset -> (file, line)
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    m = re.search(r'scan_kernel<\(int\)(\d+), \(int\)(\d+), \(bool\)(\d)>', kernel)
    tag = 'scan_kernelILi%sELi%sELb%sE' % m.groups() if m else None
    line_of, cur, infn, src_file = {}, None, False, None
    for ln in dis.splitlines():
        if ln.startswith('//---') and '.text.' in ln:
            infn = tag is not None and tag in ln
            continue
        if not infn:
            continue
        mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if mm:
            # lines of inlined header code (intrinsics) are booked under negative numbers
            cur = int(mm.group(2)) if mm.group(1).endswith('.cu') else -int(mm.group(2))
            if mm.group(1).endswith('.cu'):
                src_file = mm.group(1)
            continue
        mm = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
        if mm and cur is not None:
            line_of[int(mm.group(1), 16)] = cur
    text = {}
    if src_file and os.path.exists(src_file):
        for i, t in enumerate(open(src_file).read().splitlines(), 1):
            text[i] = t
    by_line = collections.defaultdict(collections.Counter)
    total = collections.Counter()
    for r in data:
        line = line_of.get(int(r[0], 16) - base, -1)
        d = by_line[line]
        n, ie = num(r[col['# Samples']]), num(r[col['Instructions Executed']])
        d['samples'] += n
        d['inst'] += ie
        parts = r[1].split()
        op = parts[1] if parts and parts[0].startswith('@') and len(parts) > 1 else (parts[0] if parts else '')
        if op[:4] in ('DFMA', 'DMUL', 'DADD'):
            d['fp64'] += ie
            total['fp64'] += ie
        for c in stall_cols:
            v = num(r[col[c]])
            d[c] += v
            total[c] += v
        total['samples'] += n
        total['inst'] += ie
    report(by_line, total, stall_cols, text, top)


def report(by_line, total, stall_cols, text, top):
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


def main():
    path = sys.argv[1]
    src = sys.argv[2] if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(path)))
    if not any(r and r[0] == 'Line No' for r in rows):
        return sass_only(rows, src, top)
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
