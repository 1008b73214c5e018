"""Host side of the product (no GPU): input reader, grids, neutral model, per-class selection
tables, window planners, row formatting and helper-file generators, checked against the
reference's goldens with the C oracle standing in for the CUDA kernel."""
import filecmp
import os

import numpy as np
import pytest

import util
from oracle import oracle_c, oracle_np
from ballermixplus_b200 import Grids, getConfig, getSpect, windows
from ballermixplus_b200.problem import build_problem
from ballermixplus_b200.scan import HEADER, format_rows

CASES = util.scan_cases()


# the six full-length shipped goldens are evaluated on every third row here (the GPU suite does all rows)
ROW_STRIDE = {'Example1_B1': 3, 'Example1_B2': 3, 'Example1_B2maf': 3, 'Example2_B1': 3, 'Example2_B2': 3,
              'Example2_B2maf': 3}


def host_scan_with_oracle(argv, stride=1):
    opt, data, neutral, grid, sel = util.host_objects(argv)
    prob, order = build_problem(data, neutral, sel, grid)
    with util.quiet():
        plan = windows.make_plan(data, fixSize=opt.size, r=opt.w, s=opt.step, phys=opt.phys,
                                 noCenter=opt.noCenter)
    t, lo, hi, gap = plan.arrays()
    rows = np.arange(0, len(t), stride)
    T = np.zeros(len(t)); iA = np.full(len(t), -1, np.int32); ixa = iA.copy(); ns = np.zeros(len(t), np.int32)
    T[rows], iA[rows], ixa[rows], ns[rows], _ = oracle_c.scan(prob.genpos, prob.cls, prob.G, prob.SP, prob.A,
                                                              t[rows], lo[rows], hi[rows])
    ix = np.where(ixa >= 0, ixa // prob.n_a, -1)
    ia = np.where(ixa >= 0, ixa % prob.n_a, -1)
    lines = format_rows(plan, order, T, iA, ix, ia, ns)
    keep = set(rows.tolist())
    return [HEADER] + [ln if j in keep else None for j, ln in enumerate(lines)]


@pytest.mark.parametrize('name', sorted(CASES))
def test_host_pipeline_reproduces_golden(name):
    argv, gold = CASES[name]
    lines = host_scan_with_oracle(argv, ROW_STRIDE.get(name, 1))
    # an argmax that differs from the golden must be a PROVEN tie (literal T at both grid points,
    # util.assert_ties): the two-class B1 tables make different (x, a) give T equal to ~1e-13
    tie_rows = []
    n, same, worst, ties = util.compare_scan(lines, gold, rtol=1e-9, tie_rows=tie_rows)
    assert n == sum(ln is not None for ln in lines) and n >= 4
    if tie_rows:
        util.assert_ties(argv, tie_rows)
    assert ties <= max(2, n // 50)


@pytest.mark.parametrize('name', ['Example1_B2', 'Example1_B2maf', 'Example1_B1', 'ex2_B0_s5',
                                  'Example2_B0maf_1kb-2site', 'synth_mixed_n_B2_s20'])
def test_class_tables_bit_identical_to_per_site_tables(name):
    """a4: per-class scipy evaluation == the reference's per-site evaluation, bit for bit."""
    argv, _ = CASES[name]
    opt, data, neutral, grid, sel = util.host_objects(argv)
    xs, als = grid.x[::3], grid.abeta[::5] + grid.abeta[-2:]
    per_site = oracle_np.per_site_tables(data.count, data.total, xs, als, sel.stat, data.minCount)
    for key, ref in per_site.items():
        assert np.array_equal(sel.get(*key), ref), key
    probs, logp, props = oracle_np.neutral_per_site(opt.spectfile, opt.nofreq, opt.MAF, opt.nosub,
                                                    data.count, data.total)
    assert np.array_equal(neutral.probs, probs)
    assert np.array_equal(neutral.logProbs, logp)
    assert np.array_equal(neutral.propSizes, props)


def test_helper_files_byte_identical(tmp_path):
    for name, c in util.manifest().items():
        a = c['argv']
        if '--getSpect' not in a and '--getConfig' not in a:
            continue
        infile = os.path.join(util.GOLD, a[a.index('-i') + 1])
        out = str(tmp_path / (name + '.txt'))
        with util.quiet():
            if '--getSpect' in a:
                getSpect(infile, out, '--MAF' in a, '--noSub' in a)
            else:
                getConfig(infile, out)
        assert filecmp.cmp(out, os.path.join(util.GOLD, c['output']), shallow=False), name


def test_grids_default_and_flags():
    g = Grids(None, None, False, False, None, None)
    assert len(g.x) == 10 and len(g.abeta) == 52 and len(g.A) == 31
    assert g.x[2] == 0.15000000000000002 and g.abeta[7] == 1 and isinstance(g.abeta[7], int)
    assert isinstance(g.A[0], int) and g.A[-1] == 1e8
    A, x, a = g.scan_order()
    assert (len(A), len(x), len(a)) == (31, 10, 51)            # int 5 appears twice (v1:157)
    assert A == list(set(g.A)) and x == list(set(g.x)) and a == list(set(g.abeta))
    bal = Grids(None, None, True, False, None, None)
    assert len(bal.abeta) == 45 and min(bal.abeta) == 1
    fixed = Grids('0.3', 20.0, False, False, None, '100,1000,1e4')
    assert fixed.x == [0.3] and fixed.abeta == [20.0] and fixed.A == [100.0, 1000.0, 10000.0]
    # flags the reference crashes on: documented semantics
    rng = Grids(None, None, False, False, '1000,10900,100', None)
    assert len(rng.A) == 100 and rng.A[0] == 1000.0 and rng.A[-1] == 10900.0
    assert rng.A == [float(v) for v in range(1000, 11000, 100)]
    pos = Grids(None, None, False, True, None, None)
    assert len(pos.abeta) == 7 and max(pos.abeta) == 0.8 and len(pos.x) == 9 and max(pos.x) < 1.


def test_input_reader_variants(tmp_path):
    from ballermixplus_b200 import InputData
    f = tmp_path / 'in.txt'
    f.write_text('physPos\tgenPos\tx\tn\n10.7\t1e-5\t1\t10\n20\t2e-5\t0\t10\n30\t3e-5\t7\t10\n'
                 '40\t4e-5\t10\t10\n50\t5e-5\t1\t10\n')
    with util.quiet():
        b1 = InputData(str(f), nofreq=True)
    # rows before the first non-binary count keep their value; from it on k = (k != n)  (v1:91-100)
    assert b1.count.tolist() == [1, 0, 1, 0, 1] and b1.position.tolist() == [10, 20, 30, 40, 50]
    with util.quiet():
        maf = InputData(str(f), MAF=True, phys=True, Rrate=1e-8)
    assert maf.count.tolist() == [1, 0, 3, 0, 1] and maf.minCount == 1
    assert maf.genPos.tolist() == [10.7 * 1e-8, 20 * 1e-8, 30 * 1e-8, 40 * 1e-8, 50 * 1e-8]
    with pytest.raises(SystemExit), util.quiet():
        InputData(str(f))                                        # k == 0 in DAF input (v1:60-71)
    g = tmp_path / 'in2.txt'
    g.write_text('physPos\tgenPos\tx\tn\n10\t1e-5\t3\t10\n20\t2e-5\t10\t10\n30\t3e-5\t7\t10\n')
    with util.quiet():
        b0 = InputData(str(g), nosub=True)
    assert b0.count.tolist() == [3, 7] and b0.numSites == 2 and b0.minCount == 3
    assert b0.sampSizes == {10}


def test_window_planners_match_oracle_rows():
    """a6: the vectorised/literal planners against the oracle's literal loops."""
    from ballermixplus_b200 import InputData
    rng = np.random.default_rng(7)
    pos = np.sort(rng.choice(np.arange(5, 30000), 300, replace=False))
    data = InputData.from_arrays(pos, pos * 1e-6, np.ones(300, int), np.full(300, 10))
    for kw in (dict(), dict(s=7.0), dict(r=12, s=1), dict(r=9, s=2.5), dict(r=400, s=4.0), dict(r=3, s=1.0),
               dict(fixSize=True, r=900, s=3.0), dict(fixSize=True, r=1, s=1.0), dict(fixSize=True, r=10 ** 6, s=50.0),
               dict(fixSize=True, r=2000, s=500.0, noCenter=True), dict(fixSize=True, r=300, s=150.0, noCenter=True)):
        with util.quiet():
            plan = windows.make_plan(data, phys=True, **kw)
        rows = oracle_np.scan_rows(data.position, data.genPos, data.Rrate, **kw)
        assert len(plan) == len(rows)
        for j, row in enumerate(rows):
            assert plan.gap[j] == row['gap']
            if row['gap']:
                assert plan.f0[j] == oracle_np.format_row(row, 0, 0, 0, 0, 0)
            else:
                assert (plan.t[j], plan.lo[j], plan.hi[j], plan.f0[j], plan.f1[j]) == \
                       (row['t'], row['lo'], row['hi'], row['f0'], row['f1'])


@pytest.mark.parametrize('name', ['Example1_B2', 'Example2_B2maf', 'Example1_B1', 'ex2_B0_s5',
                                  'Example2_B0maf_1kb-2site', 'synth_mixed_n_B2_s20', 'ex2_B2maf_findBal_s60'])
def test_broadcast_tables_equal_call_for_call_tables(name):
    """The one-scipy-call-per-sample-size table build against the reference's call-for-call form."""
    argv, _ = CASES[name]
    opt, data, neutral, grid, sel = util.host_objects(argv)
    for x in grid.x[::2]:
        for a in grid.abeta[::4] + grid.abeta[-3:]:
            assert np.array_equal(sel.classProbs[(x, a)], sel._slow_row(x, a)), (x, a)
