"""Shared helpers for the test-suite (not product code)."""
import contextlib
import io
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')

# the reference's seven shipped golden scans (README.md:124-172 commands)
D = 'data/'
SHIPPED = {
    'Example1_B1': ['-i', D + 'Example1_fullSweep_200kya_DAF.txt', '--spect', D + 'HC_CEU_Neut_config_for_B1.txt', '--noFreq'],
    'Example1_B2': ['-i', D + 'Example1_fullSweep_200kya_DAF.txt', '--spect', D + 'HC_CEU_Neut_DAF_spect_for_B2.txt'],
    'Example1_B2maf': ['-i', D + 'Example1_fullSweep_200kya_DAF.txt', '--spect', D + 'HC_CEU_Neut_MAF_spect_for_B2maf.txt', '--MAF'],
    'Example2_B1': ['-i', D + 'Example2_balancing_10MYA_DAF.txt', '--spect', D + 'HC_CEU_Neut_config_for_B1.txt', '--noFreq'],
    'Example2_B2': ['-i', D + 'Example2_balancing_10MYA_DAF.txt', '--spect', D + 'HC_CEU_Neut_DAF_spect_for_B2.txt'],
    'Example2_B2maf': ['-i', D + 'Example2_balancing_10MYA_DAF.txt', '--spect', D + 'HC_CEU_Neut_MAF_spect_for_B2maf.txt', '--MAF'],
    'Example2_B0maf_1kb-2site': ['-i', D + 'Example2_balancing_10MYA_MAF_nosub.txt', '--spect',
                                 D + 'HC_CEU_Neut_MAF-noSub_spect_for_B0maf.txt', '--noSub', '--MAF',
                                 '--usePhysPos', '--fixWinSize', '-w', '1000', '--step', '2'],
}


def manifest():
    with open(os.path.join(GOLD, 'manifest.json')) as fh:
        return json.load(fh)['cases']


def scan_cases():
    """name -> (argv without -o, golden file) for every scan golden (shipped + generated)."""
    out = {}
    for name, argv in SHIPPED.items():
        out[name] = (argv, os.path.join(GOLD, 'ref_out', name + '.txt'))
    for name, c in manifest().items():
        if '--getSpect' in c['argv'] or '--getConfig' in c['argv']:
            continue
        argv = list(c['argv'])
        k = argv.index('-o')
        del argv[k:k + 2]
        out[name] = (argv, os.path.join(GOLD, c['output']))
    return out


def abs_paths(argv):
    """Resolve the golden-relative file arguments."""
    out = list(argv)
    for flag in ('-i', '--spect'):
        k = out.index(flag)
        out[k + 1] = os.path.join(GOLD, out[k + 1])
    return out


def oracle_kwargs(argv):
    """argv -> keyword arguments of oracle.oracle_cli.scan_file."""
    kw = {}
    it = iter(abs_paths(argv))
    for a in it:
        if a == '-i':
            kw['infile'] = next(it)
        elif a == '--spect':
            kw['spectfile'] = next(it)
        elif a == '--noFreq':
            kw['nofreq'] = True
        elif a == '--MAF':
            kw['MAF'] = True
        elif a == '--noSub':
            kw['nosub'] = True
        elif a == '--usePhysPos':
            kw['phys'] = True
        elif a == '--rec':
            kw['Rrate'] = float(next(it))
        elif a == '--fixWinSize':
            kw['fixSize'] = True
        elif a == '-w':
            kw['w'] = int(next(it))
        elif a in ('-s', '--step'):
            kw['step'] = float(next(it))
        elif a == '--noCenter':
            kw['noCenter'] = True
        elif a == '--fixX':
            kw['x'] = next(it)
        elif a == '--fixAlpha':
            kw['abeta'] = float(next(it))
        elif a == '--findBal':
            kw['bal'] = True
        elif a == '--listA':
            kw['listA'] = next(it)
        else:
            raise ValueError(a)
    return kw


def compare_scan(lines, golden_path, rtol=1e-9, max_near_ties=0, tie_rows=None):
    """Compare output lines (None = row not evaluated) with a golden file.

    Every evaluated row must agree in the two position fields and in nSites-or-argmax as
    follows: CLR within rtol*max(|T|, 1) always; (x_hat, s_hat, A_hat, nSites) identical,
    except for at most `max_near_ties` rows whose CLR still agrees within rtol -- two grid
    points whose T differ by less than the reference's own summation noise (SURVEY.md A.4:
    the reference's sum order is unspecified), which happens with the two-class B1 tables.
    With ``tie_rows`` (a list) every such row is appended to it as (line number, got fields, golden fields)
    (line 0 = header) instead of being counted against ``max_near_ties``: the caller then PROVES each one a tie with
    ``assert_ties`` (literal T at both grid points).
    Returns (rows compared, rows byte-identical, max relative CLR difference, near ties).
    """
    with open(golden_path) as fh:
        gold = fh.read().splitlines(keepends=True)
    assert len(gold) == len(lines), f'{len(lines)} rows, golden has {len(gold)}'
    n = same = ties = 0
    worst = 0.
    for line_no, (got, ref) in enumerate(zip(lines, gold)):
        if got is None:
            continue
        n += 1
        if got == ref:
            same += 1
            continue
        a = got.rstrip('\n').split('\t')
        b = ref.rstrip('\n').split('\t')
        assert a[:2] == b[:2], f'row differs:\n got {got!r}\n ref {ref!r}'
        assert len(a) == len(b) == 7
        if 'NA' in (a[3], b[3]):
            assert a == b, f'row differs:\n got {got!r}\n ref {ref!r}'
        ta, tb = float(a[2]), float(b[2])
        rel = abs(ta - tb) / max(abs(tb), 1.)
        assert rel <= rtol, f'CLR differs by {rel:.3g}:\n got {got!r}\n ref {ref!r}'
        worst = max(worst, rel)
        if a[3:] != b[3:]:
            ties += 1
            if tie_rows is not None:
                tie_rows.append((line_no, a, b))
            else:
                assert ties <= max_near_ties, f'argmax differs:\n got {got!r}\n ref {ref!r}'
    return n, same, worst, ties


def literal_T(prob, t, lo, hi, iA, xa):
    """T of ONE grid point for one centre, the reference's formula term by term (v1:454-499) with exactly
    rounded sums (math.fsum) -> (T, nSites)."""
    import math
    g = prob.genpos
    idx = np.arange(max(int(lo), 0), min(int(hi), len(g) - 1) + 1)
    al = np.exp(-prob.A[iA] * np.abs(g[idx] - t))
    ok = (al >= 1e-8) & (g[idx] != t)
    c, a = prob.cls[idx][ok], al[ok]
    mix = a * prob.SP[xa, c] + (1. - a) * prob.G[c]
    return 2 * (math.fsum(np.log(mix)) - math.fsum(np.log(prob.G[c]))), int(ok.sum())


def assert_ties(argv, tie_rows, tol=1e-12):
    """Every row whose argmax differs from the golden must be a genuine tie: the literal T at OUR grid point and
    at the GOLDEN's grid point, for that row's centre and window, agree within tol*max(|T|, 1) -- i.e. within
    the reference's own summation noise (its sum order is set-iteration order), so which of the
    two wins is decided by rounding, not by the likelihood.  Returns the largest relative gap seen."""
    from ballermixplus_b200 import windows
    from ballermixplus_b200.problem import build_problem
    opt, data, neutral, grid, sel = host_objects(argv)
    prob, order = build_problem(data, neutral, sel, grid)
    with quiet():
        plan = windows.make_plan(data, fixSize=opt.size, r=opt.w, s=opt.step, phys=opt.phys, noCenter=opt.noCenter)
    t, lo, hi, gap = plan.arrays()
    text = {'A': [f'{v}' for v in order.A], 'x': [f'{v}' for v in order.x], 'a': [f'{v}' for v in order.a]}
    worst = 0.
    for row, got, ref in tie_rows:
        j = row - 1                                   # line 0 is the header
        vals = []
        for f in (got, ref):
            if f[3] == '0.0' and f[4] == '0.0' and f[5] == '0.0':          # the all-zero row: T = 0 by definition
                vals.append(0.)
                continue
            iA, ix, ia = text['A'].index(f[5]), text['x'].index(f[3]), text['a'].index(f[4])
            T, ns = literal_T(prob, t[j], lo[j], hi[j], iA, ix * prob.n_a + ia)
            assert ns == int(float(f[6])), f'nSites differs from the literal count in row {row}: {f}'
            vals.append(T)
        gap_rel = abs(vals[0] - vals[1]) / max(abs(vals[1]), 1.)
        assert gap_rel <= tol, f'row {row}: not a tie, literal T {vals[0]!r} vs {vals[1]!r}\n got {got}\n ref {ref}'
        worst = max(worst, gap_rel)
    return worst


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def host_objects(argv):
    """Run the product's host pipeline (no GPU) for a CLI argv -> (opt, data, neutral, grid, sel)."""
    from ballermixplus_b200 import Grids, InputData, NeutralSFS, NormalizedBetaBinom
    from ballermixplus_b200.cli import build_parser
    opt = build_parser().parse_args(abs_paths(argv))
    with quiet():
        data = InputData(opt.infile, opt.nofreq, opt.MAF, opt.nosub, opt.minCount, phys=opt.phys,
                         Rrate=opt.Rrate)
        neutral = NeutralSFS(opt.spectfile, opt.nofreq, opt.MAF, opt.nosub)
        neutral.get_neut_probs(data)
        grid = Grids(opt.x, opt.abeta, opt.bal, opt.pos, opt.seqA, opt.listA)
        sel = NormalizedBetaBinom(data, grid, opt.nofreq, opt.MAF, opt.nosub)
    return opt, data, neutral, grid, sel
