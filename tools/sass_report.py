#!/usr/bin/env python3
"""Per-kernel resource usage and SASS opcode histogram of a built library (no GPU needed).

    python tools/sass_report.py ballermixplus_b200/libblmx.so > profiles/rN_sass_report.md

Runs `cuobjdump -res-usage` (REG / STACK / SHARED per kernel) and `cuobjdump -sass`, and counts the
opcodes that matter for this path: FP64 arithmetic (DFMA, DMUL, DADD), local-memory traffic (LDL / STL =
register spills), global / shared loads and stores, bulk copies (UBLKCP = cp.async.bulk), warp shuffles /
votes / reductions and MUFU.
"""
import collections
import re
import subprocess
import sys

WATCH = ['DFMA', 'DMUL', 'DADD', 'DSETP', 'LDL', 'STL', 'LDG', 'STG', 'LDS', 'STS', 'UBLKCP', 'SYNCS', 'SHFL', 'VOTE',
         'REDUX', 'MUFU', 'ATOMG', 'RED', 'BRA', 'BAR', 'WARPSYNC']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main(path):
    res = subprocess.run(['cuobjdump', '-res-usage', path], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r'\s*Function (\S+):', line)
        if m:
            cur = m.group(1)
            continue
        if cur and 'REG:' in line:
            usage[cur] = dict(re.findall(r'(\w+):(\d+)', line))
            cur = None
    sass = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    hist = {}
    total = {}
    cur = None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            total[cur] = 0
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            hist[cur][op] += 1
    names = demangle(sorted(hist))
    print(f'# SASS report of `{path}`\n')
    print('Static instruction counts (not executed counts) per kernel, from `cuobjdump -sass`; REG / STACK / SHARED '
          'from `cuobjdump -res-usage`.  LDL / STL are local-memory (spill) instructions.\n')
    print('| kernel | REG | STACK | SHARED | instr | ' + ' | '.join(WATCH) + ' |')
    print('|---|---|---|---|---|' + '---|' * len(WATCH))
    for fn in sorted(hist, key=lambda f: names[f]):
        u = usage.get(fn, {})
        short = re.sub(r'\(anonymous namespace\)::', '', names[fn])
        short = re.sub(r'\(.*', '', short)
        row = [f'`{short}`', u.get('REG', '?'), u.get('STACK', '?'), u.get('SHARED', '?'), str(total[fn])]
        row += [str(hist[fn].get(op, 0)) for op in WATCH]
        print('| ' + ' | '.join(row) + ' |')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'ballermixplus_b200/libblmx.so')
