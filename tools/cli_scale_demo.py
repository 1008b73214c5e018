#!/usr/bin/env python3
"""End-to-end run of the drop-in CLI on a chromosome-scale file (BASELINE.json configs[3] shape):
writes a synthetic 500 k-site input, makes the helper file with --getSpect, scans every 50th site with
--usePhysPos --rec 1e-8 --rangeA 1000,10900,100, and checks a few rows against the C oracle.

    python tools/cli_scale_demo.py [--sites 500000] [--step 50]
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench                                                   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sites', type=int, default=500_000)
    ap.add_argument('--step', type=int, default=50)
    opt = ap.parse_args()
    d = tempfile.mkdtemp()
    inp, spect, out = (os.path.join(d, f) for f in ('chr.txt', 'spect.txt', 'scan.txt'))
    chrom = bench.make_chromosome(opt.sites, seed=12345)
    t0 = time.time()
    with open(inp, 'w') as fh:
        fh.write('physPos\tgenPos\tx\tn\n')
        fh.write(''.join(f'{p}\t{p * 1e-8!r}\t{k}\t{chrom["n"]}\n' for p, k in zip(chrom['pos'].tolist(), chrom['k'].tolist())))
    print(f'input written: {opt.sites} sites, {os.path.getsize(inp) / 1e6:.1f} MB, {time.time() - t0:.1f} s')
    cli = [sys.executable, '-m', 'ballermixplus_b200']
    t0 = time.time()
    subprocess.run(cli + ['-i', inp, '--getSpect', '--spect', spect], check=True, cwd=ROOT, stdout=subprocess.DEVNULL)
    print(f'--getSpect: {time.time() - t0:.1f} s')
    t0 = time.time()
    res = subprocess.run(cli + ['-i', inp, '--spect', spect, '-o', out, '--usePhysPos', '--rec', '1e-8', '--rangeA',
                                '1000,10900,100', '-s', str(opt.step)], check=True, cwd=ROOT, capture_output=True, text=True)
    wall = time.time() - t0
    rows = open(out).read().splitlines()
    n_centres = len(rows) - 1
    print(f'scan: {n_centres} centres x 51000 grid points in {wall:.1f} s wall (reader + scipy tables + GPU + writer)')
    for line in res.stdout.splitlines():
        if 'Start computing' in line or 'Scan finished' in line or 'Initializing' in line or 'Reading input' in line:
            print('   ', line.strip())
    # a few rows against the oracle
    import util
    from oracle import oracle_c
    from ballermixplus_b200.problem import build_problem
    from ballermixplus_b200.cli import build_parser
    from ballermixplus_b200 import Grids, InputData, NeutralSFS, NormalizedBetaBinom
    o = build_parser().parse_args(['-i', inp, '--spect', spect, '--usePhysPos', '--rec', '1e-8', '--rangeA', '1000,10900,100'])
    with util.quiet():
        data = InputData(o.infile, phys=True, Rrate=1e-8)
        neutral = NeutralSFS(o.spectfile, False, False, False); neutral.get_neut_probs(data)
        grid = Grids(None, None, False, False, o.seqA, None)
        sel = NormalizedBetaBinom(data, grid, False, False, False)
    prob, order = build_problem(data, neutral, sel, grid)
    pick = np.linspace(1, n_centres - 1, 4).astype(int)
    idx = (pick - 0) * opt.step
    T, iA, ixa, ns, _ = oracle_c.scan(prob.genpos, prob.cls, prob.G, prob.SP, prob.A, data.genPos[idx],
                                      np.zeros(4, np.int64), np.full(4, data.numSites - 1, np.int64))
    for j, r in zip(range(4), pick):
        f = rows[1 + r].split('\t')
        want = order.decode(T[j], iA[j], ixa[j] // prob.n_a if ixa[j] >= 0 else -1, ixa[j] % prob.n_a if ixa[j] >= 0 else -1, ns[j])
        ok = abs(float(f[2]) - float(want[0])) <= 1e-9 * max(1., abs(float(want[0]))) and f[3:] == [f'{v}' for v in want[1:]]
        print(f'    row {r}: CLI {f[2:]}  oracle {[str(v) for v in want]}  {"ok" if ok else "MISMATCH"}')
        assert ok


if __name__ == '__main__':
    main()
