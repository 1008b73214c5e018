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


def compare_scan(lines, golden_path, rtol=1e-9, max_near_ties=0):
    """Compare output lines (None = row not evaluated) with a golden file.

    Every evaluated row must agree in the two position fields and in nSites-or-argmax as
    follows: CLR within rtol*max(|T|, 1) always; (x_hat, s_hat, A_hat, nSites) identical,
    except for at most `max_near_ties` rows whose CLR still agrees within rtol -- two grid
    points whose T differ by less than the reference's own summation noise (SURVEY.md A.4:
    the reference's sum order is unspecified), which happens with the two-class B1 tables.
    Returns (rows compared, rows byte-identical, max relative CLR difference, near ties).
    """
    with open(golden_path) as fh:
        gold = fh.read().splitlines(keepends=True)
    assert len(gold) == len(lines), f'{len(lines)} rows, golden has {len(gold)}'
    n = same = ties = 0
    worst = 0.
    for got, ref in zip(lines, gold):
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
            assert ties <= max_near_ties, f'argmax differs:\n got {got!r}\n ref {ref!r}'
    return n, same, worst, ties


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
