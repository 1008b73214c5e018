"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the C ABI against
the reference's goldens and against the CPU oracle.  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

CASES = util.scan_cases()


def run_cli(argv, out):
    from ballermixplus_b200.cli import main
    with util.quiet():
        main(util.abs_paths(argv) + ['-o', out])
    with open(out) as fh:
        return fh.read().splitlines(keepends=True)


@pytest.mark.parametrize('farfield', [0, 1])
@pytest.mark.parametrize('name', sorted(CASES))
def test_cli_scan_matches_reference_golden(name, farfield, tmp_path, monkeypatch):
    """Whole runs of the drop-in CLI (all rows) against the reference's output files."""
    monkeypatch.setenv('BLMX_FARFIELD', str(farfield))
    argv, gold = CASES[name]
    lines = run_cli(argv, str(tmp_path / 'out.txt'))
    # rows whose argmax differs from the golden (two-class B1 tables: different (x, a) give the same T to ~1e-13)
    # are collected and each one is then PROVED a tie by the literal T at both grid points
    tie_rows = []
    n, same, worst, ties = util.compare_scan(lines, gold, rtol=1e-9, tie_rows=tie_rows)
    assert n == len(lines)
    gap = util.assert_ties(argv, tie_rows) if tie_rows else 0.
    assert ties <= max(2, n // 50), f'{ties} argmax differences in {n} rows'
    print(f'{name}: {n} rows, {same} byte-identical, max rel dCLR {worst:.2e}, proven ties {ties} (gap {gap:.1e})')


def _problem(name):
    from ballermixplus_b200.problem import build_problem
    argv, _ = CASES[name]
    opt, data, neutral, grid, sel = util.host_objects(argv)
    prob, order = build_problem(data, neutral, sel, grid)
    return data, prob, order


def _check_against_oracle(prob, t, lo, hi, group=None, batch=None, rtol=1e-9, farfield=None):
    from oracle import oracle_c
    from ballermixplus_b200.native import Scanner
    with Scanner(device=0, group=group, batch=batch, farfield=farfield).load(prob) as sc:
        T, iA, ix, ia, ns = sc.scan(t, lo, hi)
        pairs, _ = sc.counters()
    rT, rA, rxa, rn, rpairs = oracle_c.scan(prob.genpos, prob.cls, prob.G, prob.SP, prob.A, t, lo, hi)
    bad = np.flatnonzero((iA != rA) | (ix * prob.n_a + ia != rxa) | (ns != rn))
    bad = [j for j in bad if not (iA[j] < 0 and rA[j] < 0)]
    assert np.all(np.abs(T - rT) <= rtol * np.maximum(np.abs(rT), 1.)), np.max(np.abs(T - rT))
    assert pairs == rpairs
    return T, iA, ix, ia, ns, bad


@pytest.mark.parametrize('group,farfield', [(1, 0), (4, 0), (4, 1)])
def test_ragged_windows_and_batches(group, farfield):
    """Random, ragged, empty and out-of-range windows; small batches force several launches."""
    data, prob, order = _problem('Example2_B2')
    rng = np.random.default_rng(3)
    n = 300
    c = rng.integers(0, data.numSites, size=n)
    t = data.genPos[c].copy()
    t[::7] += 3e-7                                   # centres that are not sites
    lo = c - rng.integers(0, 500, size=n)
    hi = c + rng.integers(0, 500, size=n)
    lo[::11] = hi[::11] + 5                          # empty windows
    lo[::13] = -50                                   # clamped by the library
    hi[::17] = data.numSites + 100
    T, iA, ix, ia, ns, bad = _check_against_oracle(prob, t, lo, hi, group=group, batch=64, farfield=farfield)
    assert not bad
    empty = np.flatnonzero(lo > hi)
    assert np.all(T[empty] == 0) and np.all(iA[empty] == -1) and np.all(ns[empty] == 0)


def test_group1_and_group4_agree():
    data, prob, order = _problem('Example1_B2maf')
    from ballermixplus_b200.native import Scanner
    c = np.arange(0, data.numSites, 5)
    t, lo, hi = data.genPos[c], np.zeros(len(c), np.int64), np.full(len(c), data.numSites - 1, np.int64)
    out = []
    for g in (1, 4):
        with Scanner(device=0, group=g, farfield=0).load(prob) as sc:
            out.append(sc.scan(t, lo, hi))
    assert np.allclose(out[0][0], out[1][0], rtol=1e-12, atol=1e-12)
    for a, b in zip(out[0][1:], out[1][1:]):
        assert np.array_equal(a, b)


def test_empty_inputs_and_errors():
    from ballermixplus_b200.native import BlmxError, ScanProblem, Scanner
    data, prob, order = _problem('ex1_B2_fixgrid')
    with Scanner(device=0) as sc:
        with pytest.raises(BlmxError):
            sc.scan([0.1], [0], [1])                 # scan before load
        sc.load(prob)
        T, iA, ix, ia, ns = sc.scan([], [], [])      # empty batch
        assert len(T) == 0
        with pytest.raises(BlmxError):
            sc.set_option('group', 3)
        bad = ScanProblem(prob.genpos, np.full_like(prob.cls, 10 ** 6), prob.G, prob.SP, prob.A, prob.n_x, prob.n_a)
        with pytest.raises(BlmxError):
            sc.load(bad)
    # a problem with no sites at all
    empty = ScanProblem(np.zeros(0), np.zeros(0, np.int32), prob.G, prob.SP, prob.A, prob.n_x, prob.n_a)
    with Scanner(device=0).load(empty) as sc:
        T, iA, ix, ia, ns = sc.scan([0.5, 1.0], [0, 0], [10, -1])
        assert np.all(T == 0) and np.all(iA == -1) and np.all(ns == 0)


def test_unsorted_genpos_and_duplicates():
    """The reference never checks that the input is sorted; distance pruning must switch off."""
    data, prob, order = _problem('synth_dup_genpos_B2')
    from ballermixplus_b200.native import ScanProblem
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(prob.genpos))
    shuffled = ScanProblem(prob.genpos[perm], prob.cls[perm], prob.G, prob.SP, prob.A, prob.n_x, prob.n_a)
    n = len(perm)
    t = prob.genpos[rng.integers(0, n, size=40)]
    lo = rng.integers(0, n // 2, size=40)
    hi = lo + rng.integers(0, n // 2, size=40)
    for p in (prob, shuffled):
        for ff in (0, 1):
            T, iA, ix, ia, ns, bad = _check_against_oracle(p, t, lo, hi, farfield=ff)
            assert not bad


def test_large_xa_grid_needs_several_passes():
    """n_x * n_a > 512: the kernel sweeps the (x, a) grid in blocks of 512."""
    from ballermixplus_b200.native import ScanProblem
    data, prob, order = _problem('ex2_B0_s5')
    reps = 3                                          # 510 * 3 = 1530 grid points
    SP = np.concatenate([prob.SP * (1 + 1e-3 * r) for r in range(reps)])
    big = ScanProblem(prob.genpos, prob.cls, prob.G, SP, prob.A, prob.n_x * reps, prob.n_a)
    c = np.arange(0, data.numSites, 9)
    t, lo, hi = data.genPos[c], np.zeros(len(c), np.int64), np.full(len(c), data.numSites - 1, np.int64)
    for ff in (0, 1):
        T, iA, ix, ia, ns, bad = _check_against_oracle(big, t, lo, hi, farfield=ff)
        assert not bad


def test_extreme_tables_stay_in_range():
    """Probabilities that underflow, huge likelihood ratios and alpha -> 1: the exponent
    bookkeeping must keep every product in range (T up to several thousand)."""
    from ballermixplus_b200.native import ScanProblem
    rng = np.random.default_rng(11)
    n_sites, C, n_xa = 4000, 6, 40
    g = np.sort(rng.random(n_sites)) * 1e-3
    g[100:104] = g[100] + np.array([0, 1e-15, 2e-15, 1e-13])      # alpha within 1e-11 of 1
    cls = rng.integers(0, C, size=n_sites).astype(np.int32)
    G = np.array([0.5, 0.2, 0.1, 0.1, 0.05, 0.05])
    SP = rng.random((n_xa, C)) * G * 2
    SP[3] = G * 50.                       # every factor up to 50: T ~ 2*n*log(50) overflows a double product
    SP[5, :] = G * 1e-300                 # D = -1 (+1e-300)
    SP[6, 2] = 0.                         # exact zero: T = -inf for windows touching class 2 at alpha = 1
    SP[7] = G * 1e12
    prob = ScanProblem(g, cls, G, SP, [50., 2000., 1e5, 1e7], 8, 5)
    c = np.array([0, 100, 101, 103, 2000, 3999])
    t = g[c]
    lo, hi = np.zeros(len(c), np.int64), np.full(len(c), n_sites - 1, np.int64)
    for group, ff in ((1, 0), (4, 0), (4, 1)):
        T, iA, ix, ia, ns, bad = _check_against_oracle(prob, t, lo, hi, group=group, farfield=ff)
        assert not bad
        assert np.all(np.isfinite(T)) and T.max() > 2000.


def test_synthetic_n200_against_oracle_sample():
    """cfg4-shaped synthetic input (n = 200, 200 classes, 100 A x 510 grid) at a size the C oracle
    finishes in seconds, plus size-independent properties on the full centre list."""
    import bench
    from ballermixplus_b200.native import Scanner
    chrom = bench.make_chromosome(20000, seed=12345)
    prob, grid_shape = bench.make_problem([chrom])[0], None
    n = len(prob.genpos)
    rng = np.random.default_rng(2)
    c = np.sort(rng.choice(n, size=24, replace=False))
    t, lo, hi = prob.genpos[c], np.zeros(len(c), np.int64), np.full(len(c), n - 1, np.int64)
    T, iA, ix, ia, ns, bad = _check_against_oracle(prob, t, lo, hi, farfield=0)
    assert not bad
    Tf, *_rest, badf = _check_against_oracle(prob, t, lo, hi, farfield=1)
    assert not badf
    assert np.max(np.abs(Tf - T)) <= 1e-11 * max(1., np.max(np.abs(T)))
    # properties at full size: all centres, (1) group 1 == group 4, (2) results do not depend on batching,
    # (3) a window cut at the alpha reach of the smallest A changes nothing
    call = np.arange(0, n, 40)
    ta, loa, hia = prob.genpos[call], np.zeros(len(call), np.int64), np.full(len(call), n - 1, np.int64)
    with Scanner(device=0, group=4, farfield=0).load(prob) as sc:
        r4 = sc.scan(ta, loa, hia)
        reach = 18.420680743952367 / prob.A.min() * 1.01
        lo2 = np.searchsorted(prob.genpos, ta - reach, 'left')
        hi2 = np.searchsorted(prob.genpos, ta + reach, 'right') - 1
        rcut = sc.scan(ta, lo2, hi2)
        sc.set_option('batch', 100)
        rb = sc.scan(ta, loa, hia)
    with Scanner(device=0, group=1).load(prob) as sc:
        r1 = sc.scan(ta, loa, hia)
    with Scanner(device=0, group=4, farfield=1).load(prob) as sc:
        rf = sc.scan(ta, loa, hia)
        far = sc.counters_all()
    assert far['far_sites'] > 0.4 * far['pairs'] and far['far_terms'] > 0      # the far field was actually used
    assert np.allclose(r4[0], rf[0], rtol=1e-11, atol=1e-11)
    for a, b in zip(r4[1:], rf[1:]):
        assert np.array_equal(a, b)
    for other in (rcut, rb):
        for a, b in zip(r4, other):
            assert np.array_equal(a, b)
    assert np.allclose(r4[0], r1[0], rtol=1e-11, atol=1e-11)
    for a, b in zip(r4[1:], r1[1:]):
        assert np.array_equal(a, b)


def test_default_grid_with_windows_of_100k_sites():
    """The reference's default A grid starts at A = 100 (v1:162): on a 150 k-site chromosome such a window holds
    ~130 k sites and the far field runs over hundreds of superblocks; at the other end A = 1e8 leaves a window of a
    few sites.  24 centres (ends included), all kernel modes' far/direct agreement, the oracle on every one."""
    import bench
    from ballermixplus_b200.native import Scanner
    chrom = bench.make_chromosome(150000, seed=77)
    prob = bench.make_problem([chrom], range_a=None)[0]
    assert len(prob.A) == 31 and prob.A.min() == 100.0
    n = len(prob.genpos)
    c = np.concatenate(([0, 1, n - 2, n - 1], np.linspace(5, n - 6, 20).astype(np.int64)))
    t, lo, hi = prob.genpos[c], np.zeros(len(c), np.int64), np.full(len(c), n - 1, np.int64)
    mid = prob.genpos[n // 2]
    assert np.count_nonzero(np.exp(-100.0 * np.abs(prob.genpos - mid)) >= 1e-8) > 100_000     # the A = 100 window
    T, iA, ix, ia, ns, bad = _check_against_oracle(prob, t, lo, hi, farfield=1)      # CLR, site pairs: asserted inside
    assert not bad, (bad, iA[bad], ns[bad])
    with Scanner(device=0, farfield=0).load(prob) as sc:
        direct = sc.scan(t, lo, hi)
    assert np.all(np.abs(T - direct[0]) <= 1e-10 * np.maximum(np.abs(direct[0]), 1.))
    for a, b in zip((iA, ix, ia, ns), direct[1:]):
        assert np.array_equal(a, b)


def test_report_all_reports_the_best_grid_point_whatever_its_sign():
    """Option report_all (the maximum starts from -inf instead of the reference's 0, v1:451): on neutral
    synthetic data most centres have no grid point with T > 0, so the ordinary rows are all-zero and compare
    nothing; with report_all every centre carries a real (negative) T, argmax and nSites, checked against
    the oracle run the same way."""
    import bench
    from oracle import oracle_c
    from ballermixplus_b200.native import Scanner
    chrom = bench.make_chromosome(30000, seed=4)
    prob = bench.make_problem([chrom])[0]
    n = len(prob.genpos)
    c = np.linspace(0, n - 1, 40).astype(np.int64)
    t, lo, hi = prob.genpos[c], np.zeros(len(c), np.int64), np.full(len(c), n - 1, np.int64)
    lo[3], hi[3] = 50, 40                                     # an empty window stays the all-zero row
    rT, rA, rxa, rn, _ = oracle_c.scan(prob.genpos, prob.cls, prob.G, prob.SP, prob.A, t, lo, hi, report_all=True)
    assert np.count_nonzero(rT < 0) > 20 and rA[3] == -1
    for ff in (0, 1):
        with Scanner(device=0, farfield=ff) as sc:
            sc.set_option('report_all', 1)
            sc.load(prob)
            T, iA, ix, ia, ns = sc.scan(t, lo, hi)
            sc.set_option('report_all', 0)
            T0, iA0, *_ = sc.scan(t, lo, hi)
        xa = np.where(iA >= 0, ix * prob.n_a + ia, -1)
        assert np.array_equal(iA, rA) and np.array_equal(xa, rxa) and np.array_equal(ns, rn)
        assert np.all(np.abs(T - rT) <= 1e-9 * np.maximum(np.abs(rT), 1.))
        assert np.all(T0 >= 0) and np.all((iA0 >= 0) == (rT > 0))          # the reference's rule again


def test_thousands_of_classes():
    """Inputs with many sample sizes have thousands of (k, n) classes (the reference supports them,
    v1:331-355): 2 400 classes here, most of them with no site in any one window.  Parity with the oracle in
    all kernel modes, and the time per (centre, A) item for the record."""
    from ballermixplus_b200.native import Scanner, ScanProblem
    rng = np.random.default_rng(23)
    n_sites, C, n_x, n_a = 60000, 2400, 5, 21
    g = np.sort(rng.random(n_sites)) * 0.05
    w = rng.random(C) ** 4 + 1e-4
    w[:3] += 2.0                                               # three big classes, a long tail of rare ones
    cls = rng.choice(C, size=n_sites, p=w / w.sum()).astype(np.int32)
    G = rng.random(C) * 0.1 + 1e-3
    SP = G[None, :] * 10.0 ** rng.normal(0, 0.4, (n_x * n_a, C))
    A = np.array([300., 1000., 3000., 1e4, 1e5])
    prob = ScanProblem(g, cls, G, SP, A, n_x, n_a)
    c = rng.integers(0, n_sites, 48)
    t, lo, hi = g[c], np.zeros(48, np.int64), np.full(48, n_sites - 1, np.int64)
    for group, ff in ((1, 0), (4, 0), (4, 1)):
        T, iA, ix, ia, ns, bad = _check_against_oracle(prob, t, lo, hi, group=group, farfield=ff)
        assert not bad
    call = np.arange(0, n_sites, 30)
    with Scanner(device=0).load(prob) as sc:
        sc.set_option('timing', 1)
        sc.scan(g[call], np.zeros(len(call), np.int64), np.full(len(call), n_sites - 1, np.int64))
        ms, _ = sc.kernel_ms()
        cnt = sc.counters_all()
    print(f'{C} classes: {len(call) * len(A)} items in {ms:.1f} ms = {1e3 * ms / (len(call) * len(A)):.1f} us/item, '
          f'{cnt["pairs"] / (len(call) * len(A)):.0f} sites per item')


def test_cli_under_torchrun_nccl(tmp_path):
    """`torchrun --nproc-per-node 2 -m ballermixplus_b200 ...` with one GPU per rank: rows go from the scan
    kernel to the NCCL gather without leaving the device; rank 0 writes the same file.  Needs two GPUs."""
    import subprocess
    import sys
    from ballermixplus_b200 import native
    if native.device_count() < 2:
        pytest.skip('needs at least two GPUs')
    argv, gold = CASES['Example2_B2']
    out = str(tmp_path / 'out.txt')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29541', '-m', 'ballermixplus_b200'] + util.abs_paths(argv) + ['-o', out]
    res = subprocess.run(cmd, cwd=util.ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    with open(out) as fh:
        lines = fh.read().splitlines(keepends=True)
    n, same, worst, ties = util.compare_scan(lines, gold)
    assert n == len(lines)


def _cli_vs_oracle(argv, tmp_path, prove_ties=False):
    """Run the CLI on the GPU and the same host pipeline with the C oracle as the kernel."""
    from oracle import oracle_c
    from ballermixplus_b200 import windows
    from ballermixplus_b200.problem import build_problem
    from ballermixplus_b200.scan import HEADER, format_rows
    lines = run_cli(argv, str(tmp_path / 'out.txt'))
    opt, data, neutral, grid, sel = util.host_objects(argv)
    prob, order = build_problem(data, neutral, sel, grid)
    with util.quiet():
        plan = windows.make_plan(data, fixSize=opt.size, r=opt.w, s=opt.step, phys=opt.phys, noCenter=opt.noCenter)
    t, lo, hi, gap = plan.arrays()
    T, iA, ixa, ns, _ = oracle_c.scan(prob.genpos, prob.cls, prob.G, prob.SP, prob.A, t, lo, hi)
    ix = np.where(ixa >= 0, ixa // prob.n_a, -1)
    ia = np.where(ixa >= 0, ixa % prob.n_a, -1)
    ref_path = str(tmp_path / 'oracle.txt')
    with open(ref_path, 'w') as fh:
        fh.write(HEADER)
        fh.writelines(format_rows(plan, order, T, iA, ix, ia, ns))
    tie_rows = [] if prove_ties else None
    res = util.compare_scan(lines, ref_path, rtol=1e-9, tie_rows=tie_rows)
    if tie_rows:
        util.assert_ties(argv, tie_rows)
    return lines, res


D = 'data/'


def test_flags_the_reference_crashes_on(tmp_path):
    """--rangeA, --findPos and --minCount with --noFreq have no reference output:
    checked against the oracle driven by the same host objects, and --rangeA against --listA."""
    ex1 = ['-i', D + 'Example1_fullSweep_200kya_DAF.txt', '--spect', D + 'HC_CEU_Neut_DAF_spect_for_B2.txt']
    a, _ = _cli_vs_oracle(ex1 + ['--rangeA', '500,4500,1000', '-s', '40'], tmp_path)
    b, _ = _cli_vs_oracle(ex1 + ['--listA', '500,1500,2500,3500,4500', '-s', '40'], tmp_path)
    assert a == b and len(a) == 20
    # the --findPos x grid holds both x and 1-x, whose folded tables are mathematically identical:
    # every row is an exact tie between the two, resolved by rounding: each differing row is proved a tie
    lines, (n, same, worst, ties) = _cli_vs_oracle(ex1 + ['--findPos', '-s', '60'], tmp_path, prove_ties=True)
    assert n == 14 and all(float(l.split('\t')[3]) < 1.0 for l in lines[1:])
    b1 = ['-i', D + 'Example2_balancing_10MYA_DAF.txt', '--spect', D + 'HC_CEU_Neut_config_for_B1.txt', '--noFreq']
    _cli_vs_oracle(b1 + ['--minCount', '2', '-s', '70', '--listA', '100,1000,1e4'], tmp_path, prove_ties=True)


def test_calcballer_drop_in_one_centre_at_a_time():
    """The reference's per-centre seam (v1:436): same signature, same 5-element list."""
    from oracle import oracle_np
    from ballermixplus_b200 import calcBaller
    argv, _ = CASES['ex1_B2_w15_s7p5']
    opt, data, neutral, grid, sel = util.host_objects(argv)
    A, x, a = grid.scan_order()
    for i in (0, 233, 756):
        lo, hi = max(0, i - 15), min(data.numSites - 1, i + 16)
        got = calcBaller(np.arange(lo, hi + 1), data.genPos[i], data, neutral, sel, grid)
        want = oracle_np.calc_baller(lo, hi, data.genPos[i], data.genPos, neutral.probs, neutral.logProbs,
                                     neutral.propSizes, {k: sel.get(*k) for k in sel.classProbs}, A, x, a)
        assert got[1:] == want[1:] and type(got[2]) is type(want[2]) and type(got[3]) is type(want[3])
        assert abs(got[0] - want[0]) <= 1e-9 * max(1., abs(want[0]))
    assert calcBaller(np.arange(0), data.genPos[5], data, neutral, sel, grid) == [0., 0., 0., 0., 0.]


def test_range_checked_build_sees_no_violation():
    """The -DBLMX_CHECKED build counts every out-of-range index the scan kernel would form
    (compute-sanitizer is closed on the GPU pool): ragged windows, shuffled input, tiny and
    huge grids, both kernel modes -- the counter must stay 0 and results must not change."""
    import subprocess
    import sys
    lib = os.path.join(util.ROOT, 'ballermixplus_b200', 'libblmx_checked.so')
    assert os.path.exists(lib), 'run __graft_entry__.build() first'
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import util
from ballermixplus_b200.native import Scanner, ScanProblem
from ballermixplus_b200.problem import build_problem
rng = np.random.default_rng(17)
bad = 0
for name in ('Example2_B2', 'Example1_B1', 'ex1_B2_fixgrid', 'synth_mixed_n_B2_s20', 'synth_n200_B2_s700'):
    argv, _ = util.scan_cases()[name]
    opt, data, neutral, grid, sel = util.host_objects(argv)
    prob, order = build_problem(data, neutral, sel, grid)
    n = data.numSites
    perm = rng.permutation(n)
    shuffled = ScanProblem(prob.genpos[perm], prob.cls[perm], prob.G, prob.SP, prob.A, prob.n_x, prob.n_a)
    c = rng.integers(0, n, size=150)
    t = data.genPos[c] + np.where(rng.random(150) < 0.2, 1e-7, 0.0)
    lo = c - rng.integers(-20, 900, size=150)
    hi = c + rng.integers(-20, 900, size=150)
    for p in (prob, shuffled):
        res = []
        for ff in (0, 1):
            with Scanner(device=0, farfield=ff, batch=37).load(p) as sc:
                res.append(sc.scan(t, lo, hi))
                bad += sc.counters_all()['range_violations']
        assert np.allclose(res[0][0], res[1][0], rtol=1e-10, atol=1e-10)
print('VIOLATIONS', bad)
''' % (util.ROOT, os.path.join(util.ROOT, 'tests'))
    env = dict(os.environ, BLMX_LIB=lib)
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'VIOLATIONS 0' in out.stdout, out.stdout[-500:]


def test_full_size_chromosome_properties():
    """BASELINE.json's full size (chromosome 1 of the 10 M-site genome: 846 k sites, n = 200,
    100 A x 510 grid, windows of up to 53 k sites): the two kernel modes agree on 160 centres, the
    C oracle confirms >= 16 centres whose row carries a real maximum (T > 0) plus 8 more with option
    report_all (a real T on every one of them), and centres give the same row whatever batch they travel in."""
    import bench
    from oracle import oracle_c
    from ballermixplus_b200.native import Scanner
    sizes = bench.genome_sizes(10_000_000)
    chrom = bench.make_chromosome(int(sizes[0]), seed=12345)
    prob = bench.make_problem([chrom])[0]
    n = len(prob.genpos)
    assert n > 800_000
    c = np.linspace(0, n - 1, 160).astype(np.int64)
    t, lo, hi = prob.genpos[c], np.zeros(len(c), np.int64), np.full(len(c), n - 1, np.int64)
    with Scanner(device=0, farfield=1).load(prob) as sc:
        far = sc.scan(t, lo, hi)
        cnt = sc.counters_all()
        sc.set_option('batch', 7)
        far_b = sc.scan(t, lo, hi)
    with Scanner(device=0, farfield=0).load(prob) as sc:
        direct = sc.scan(t, lo, hi)
    assert cnt['far_sites'] > 0.6 * cnt['pairs']
    for a, b in zip(far, far_b):
        assert np.array_equal(a, b)
    assert np.all(np.abs(far[0] - direct[0]) <= 1e-10 * np.maximum(np.abs(direct[0]), 1.))
    for a, b in zip(far[1:], direct[1:]):
        assert np.array_equal(a, b)
    assert cnt['pairs'] > 160 * 100 * 10_000             # ~12.7 k sites per (centre, A) on average
    # the oracle on centres whose row carries a real maximum (neutral data: about a quarter of them) ...
    winners = np.flatnonzero(far[1] >= 0)
    assert len(winners) >= 16, len(winners)
    pick = winners[np.linspace(0, len(winners) - 1, 20).astype(np.int64)]
    rT, rA, rxa, rn, _ = oracle_c.scan(prob.genpos, prob.cls, prob.G, prob.SP, prob.A, t[pick], lo[pick], hi[pick])
    got_xa = np.where(far[1][pick] >= 0, far[2][pick] * prob.n_a + far[3][pick], -1)
    assert np.count_nonzero(rA >= 0) >= 16
    assert np.array_equal(far[1][pick], rA) and np.array_equal(got_xa, rxa)
    assert np.array_equal(far[4][pick], rn)
    assert np.all(np.abs(far[0][pick] - rT) <= 1e-9 * np.maximum(np.abs(rT), 1.))
    # ... and, with report_all on both sides, on centres whose ordinary row is all-zero
    losers = np.flatnonzero(far[1] < 0)[:8]
    rT, rA, rxa, rn, _ = oracle_c.scan(prob.genpos, prob.cls, prob.G, prob.SP, prob.A, t[losers], lo[losers],
                                       hi[losers], report_all=True)
    with Scanner(device=0, farfield=1) as sc:
        sc.set_option('report_all', 1)
        sc.load(prob)
        aT, aA, ax, aa, an = sc.scan(t[losers], lo[losers], hi[losers])
    assert np.all(rA >= 0) and np.all(rT <= 0)
    assert np.array_equal(aA, rA) and np.array_equal(ax * prob.n_a + aa, rxa) and np.array_equal(an, rn)
    assert np.all(np.abs(aT - rT) <= 1e-9 * np.maximum(np.abs(rT), 1.))


def test_randomised_parity_campaign():
    """20 s of tools/fuzz_parity.py: random tables (huge, underflowing, exact zeros), grids, windows,
    sorted and shuffled inputs, all three kernel modes against the C oracle."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(util.ROOT, 'tools', 'fuzz_parity.py'), '--seconds', '20',
                          '--seed', '7'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    assert 'DONE:' in out.stdout


def test_cli_under_torchrun_two_ranks(tmp_path):
    """`torchrun --nproc-per-node 2 -m ballermixplus_b200 ...`: the centres are sharded over two ranks
    (sharing the one GPU of the test box, so the gather runs over gloo), rank 0 writes the same file."""
    import subprocess
    import sys
    argv, gold = CASES['ex1_B2_w20_s10']
    out = str(tmp_path / 'out.txt')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29533', '-m', 'ballermixplus_b200'] + util.abs_paths(argv) + ['-o', out]
    res = subprocess.run(cmd, cwd=util.ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    with open(out) as fh:
        lines = fh.read().splitlines(keepends=True)
    n, same, worst, ties = util.compare_scan(lines, gold)
    assert n == len(lines)


def test_c_abi_multi_gpu_entry():
    """blmx_scan_sharded (include/blmx_mgpu.h): one thread per GPU, NCCL communicator from ncclCommInitAll, rows
    on rank 0 identical to a single-GPU scan.  Runs tools/mgpu_abi_demo.py in a process without torch."""
    import subprocess
    import sys
    from ballermixplus_b200 import native
    if native.device_count() < 2:
        pytest.skip('needs at least two GPUs')
    if not os.path.exists(os.path.join(util.ROOT, 'ballermixplus_b200', 'libblmx_mgpu.so')):
        pytest.skip('libblmx_mgpu.so not built')
    res = subprocess.run([sys.executable, os.path.join(util.ROOT, 'tools', 'mgpu_abi_demo.py'), '--gpus', '2'],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and 'OK:' in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]


def test_oneshot_and_timing_entry_points():
    """blmx_scan_oneshot (create + load + scan + destroy) and the per-launch kernel timing."""
    import ctypes as C
    from ballermixplus_b200 import native
    data, prob, order = _problem('ex1_B2_fixgrid')
    c = np.arange(0, data.numSites, 7)
    t = np.ascontiguousarray(data.genPos[c])
    lo = np.zeros(len(c), np.int64)
    hi = np.full(len(c), data.numSites - 1, np.int64)
    with native.Scanner(device=0).load(prob) as sc:
        sc.set_option('timing', 1)
        want = sc.scan(t, lo, hi)
        ms, n_launch = sc.kernel_ms()
        assert n_launch == 1 and 0 < ms < 1000
    T = np.zeros(len(c)); iA, ix, ia, ns = (np.zeros(len(c), np.int32) for _ in range(4))
    res = native.Result(*(a.ctypes.data_as(C.c_void_p) for a in (T, iA, ix, ia, ns)))
    st = prob.as_struct()
    rc = native.lib().blmx_scan_oneshot(0, C.byref(st), len(c), t.ctypes.data_as(C.c_void_p),
                                        lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), C.byref(res))
    assert rc == 0, native.lib().blmx_last_error()
    for a, b in zip((T, iA, ix, ia, ns), want):
        assert np.array_equal(a, b)
    # errors come back as codes with a message, never as exceptions or exits
    rc = native.lib().blmx_scan_oneshot(0, C.byref(st), len(c), None, None, None, C.byref(res))
    assert rc == -1 and b'null' in native.lib().blmx_last_error()
