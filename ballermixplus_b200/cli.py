"""Command line -- SURVEY.md §8 rows a8/b: the flag surface of BalLeRMix+_v1.py.

Same flags, defaults, precedence and console messages as the reference's ``main``
(/root/reference/BalLeRMix+_v1.py:715-802); one extra flag, ``--device``, picks the
GPU.  Differences, all on paths where the reference raises (SURVEY.md appendix A.2,
DESIGN.md "flags without a reference behaviour"): ``--rangeA``, ``--findPos`` and
``--minCount`` with ``--noFreq`` work here.
"""
import argparse
import os
import sys
from datetime import datetime

from .grids import Grids
from .helpers import getConfig, getSpect
from .inputs import InputData
from .neutral import NeutralSFS
from .scan import Scan
from .selection import NormalizedBetaBinom


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('-i', '--input', dest='infile', required=True, help='Path and name of your input file.\n')
    p.add_argument('-o', '--output', dest='outfile', help='Path and name of your output file.\n')
    p.add_argument('--spect', dest='spectfile', required=True,
                   help='Path and name of the allele frequency spectrum file or configuration file.\n')
    p.add_argument('--minCount', dest='minCount', default=1,
                   help='If rare variants are removed from the input, please provide the smallest '
                        'allele count included in the input. Default value is 1.')
    p.add_argument('--getSpect', dest='getSpec', action='store_true', default=False,
                   help='Generate the frequency spectrum file from the concatenated input (-i) into --spect.')
    p.add_argument('--getConfig', dest='getConfig', action='store_true', default=False,
                   help='Generate the substitution/polymorphism configuration file from the '
                        'concatenated input (-i) into --spect.')
    p.add_argument('--findBal', dest='bal', action='store_true', default=False,
                   help='Only look for footprints of balancing selection.')
    p.add_argument('--findPos', dest='pos', action='store_true', default=False,
                   help='Only look for footprints of positive selection.')
    p.add_argument('--noFreq', dest='nofreq', action='store_true', default=False,
                   help='Compute B_1 (ignore allele frequencies).')
    p.add_argument('--noSub', dest='nosub', action='store_true', default=False,
                   help='Do not include substitutions: B_0 or B_0maf.')
    p.add_argument('--MAF', dest='MAF', action='store_true', default=False,
                   help='Use minor allele frequencies: B_2maf or B_0maf.')
    p.add_argument('--usePhysPos', action='store_true', dest='phys', default=False,
                   help='Use physical positions (times --rec) instead of genetic positions.')
    p.add_argument('--rec', dest='Rrate', default=1e-6, type=float,
                   help='Uniform recombination rate in cM/nt (default 1e-6).')
    p.add_argument('--fixWinSize', action='store_true', dest='size', default=False,
                   help='Fix the size (nt) of the sliding windows; give the width with -w.')
    p.add_argument('-w', '--window', dest='w', type=int, default=0,
                   help='Sites flanking the test locus on either side, or the window width in bp with --fixWinSize.')
    p.add_argument('--noCenter', action='store_true', dest='noCenter', default=False,
                   help='Windows not centred on informative sites (needs --fixWinSize -w and --usePhysPos).')
    p.add_argument('-s', '--step', dest='step', type=float, default=1,
                   help='Step size in bp (--noCenter) or in informative sites. Default 1.')
    p.add_argument('--fixX', dest='x', help='Fix the presumed equilibrium frequency.')
    p.add_argument('--fixAlpha', dest='abeta', type=float, default=None,
                   help='Fix the alpha parameter of the beta-binomial distribution.')
    p.add_argument('--rangeA', dest='seqA', help='<Amin>,<Amax>,<Astep> grid for the linkage parameter A.')
    p.add_argument('--listA', dest='listA', help='Comma-separated list of A values.')
    p.add_argument('--device', dest='device', type=int, default=0, help='CUDA device to run the scan on.')
    return p


def main(argv=None):
    argv = sys.argv[1:] if argv is None else list(argv)
    if int(os.environ.get('RANK', '0')) != 0:          # torchrun: only rank 0 talks and writes
        sys.stdout = open(os.devnull, 'w')
    parser = build_parser()
    if len(argv) == 0:
        parser.print_help()
        sys.exit()
    opt = parser.parse_args(argv)

    if opt.getSpec:
        print('You\'ve chosen to generate site frequency spectrum...')
        print('Concatenated input: %s \nSpectrum file: %s' % (opt.infile, opt.spectfile))
        getSpect(opt.infile, opt.spectfile, opt.MAF, opt.nosub)
        sys.exit()
    elif opt.getConfig:
        print('You\'ve chosen to generate the substitution-polymorphism configuration...')
        print('Concatenated input: %s \nConfiguration file: %s' % (opt.infile, opt.spectfile))
        getConfig(opt.infile, opt.spectfile)
        sys.exit()

    print(f'\n{datetime.now()}. Reading input from {opt.infile}')
    data = InputData(opt.infile, opt.nofreq, opt.MAF, opt.nosub, opt.minCount, phys=opt.phys,
                     Rrate=opt.Rrate)
    Neutral = NeutralSFS(opt.spectfile, opt.nofreq, opt.MAF, opt.nosub)

    print(f'\n{datetime.now()}. Initializing...')
    print('Retrieving per-site neutral probabilities...')
    Neutral.get_neut_probs(data)

    grid = Grids(opt.x, opt.abeta, opt.bal, opt.pos, opt.seqA, opt.listA)
    print('\nOptimizing over x= ' + ', '.join(['%g' % (x) for x in grid.x]))
    print('\n \t alpha= ' + ', '.join([str(a) for a in grid.abeta]))
    print('\n \t A= ' + ', '.join([str(A) for A in grid.A]))

    Sel_Probs = NormalizedBetaBinom(data, grid, opt.nofreq, opt.MAF, opt.nosub)

    print('\n%s. Start computing likelihood raito...' % (datetime.now()))
    Scan(data, Neutral, Sel_Probs, grid, opt.outfile, fixSize=opt.size, r=opt.w, s=opt.step,
         phys=opt.phys, noCenter=opt.noCenter, device=opt.device)
    print(f'\n{datetime.now()}. Pipeline finished.')


if __name__ == '__main__':
    main()
