"""CPU ORACLE (test infrastructure, not product code) -- end-to-end driver.

Restates the reference's scan pipeline (/root/reference/BalLeRMix+_v1.py:715-802,
"v1") on top of oracle_np: read input (v1:8-131), read helper file (v1:180-304),
build grids (v1:134-175), per-site selection tables (v1:319-433), window modes
(v1:510-640) and row formatting, so that whole output files can be compared with
the reference's goldens.  PARITY PINNED by tests/test_oracle_golden.py.

``centre_stride`` evaluates only every k-th output row (the others are returned
as None) so that the CPU test-suite stays within minutes.
"""
import numpy as np

from . import oracle_np as onp

HEADER = 'physPos\tgenPos\tCLR\tx_hat\ts_hat\tA_hat\tnSites\n'


def read_input(infile, nofreq, MAF, nosub, minCount, phys, Rrate):
    """v1:8-131, per line."""
    pos, gen, cnt, tot = [], [], [], []
    translate = False
    pt = 1 - int(phys)
    with open(infile) as fh:
        next(fh)
        for line in fh:
            f = line.strip().split('\t')
            p, k, n = int(float(f[0])), int(f[2]), int(f[3])
            if nofreq:                                     # v1:91-100
                if not translate:
                    if k not in (0, 1):
                        translate = True
                        k = int(k != n)
                else:
                    k = int(k != n)
            pos.append(p); cnt.append(k); tot.append(n)
            gen.append(float(f[pt]) * (1 - pt) * Rrate + float(f[pt]) * pt)   # v1:103,124
    pos = np.array(pos); gen = np.array(gen); cnt = np.array(cnt); tot = np.array(tot)
    if not nofreq:
        if nosub and np.sum(cnt == tot) > 0:              # v1:41-50
            keep = np.where(cnt != tot)
            pos, gen, cnt, tot = pos[keep], gen[keep], cnt[keep], tot[keep]
        if MAF:                                            # v1:53-59
            cnt = np.where(cnt > tot / 2, tot - cnt, cnt)
            minCount = min(k for k in cnt if k > 0)
        else:                                              # v1:60-74
            assert np.sum(cnt == 0) == 0
            minCount = min(cnt)
    else:
        minCount = int(minCount)
    return pos, gen, cnt, tot, minCount


def build_grids(x=None, abeta=None, bal=False, listA=None):
    """v1:136-175 (without the two crashing branches)."""
    xs = [float(x)] if x is not None else [.05 * i for i in range(1, 11)]
    full = ([0.001, 0.01, 0.05, 0.1, 0.2, 0.5, 0.8] + [i for i in range(1, 10)]
            + [5 * i for i in range(1, 20)] + [10 * i for i in range(10, 21)]
            + [300, 500, 1e3, 1e4, 1e6, 1e9])
    if abeta is not None:
        al = [float(abeta)]
    elif bal:
        al = full[7:]
    else:
        al = full
    if listA:
        As = [float(v) for v in listA.split(',')]
    else:
        As = ([100 * i for i in range(1, 12)] + [200 * i for i in range(6, 13)]
              + [500 * i for i in range(5, 10)] + [1000 * i for i in range(5, 11)] + [1e6, 1e8])
    return xs, al, As


def scan_file(infile, spectfile, nofreq=False, MAF=False, nosub=False, minCount=1, phys=False,
              Rrate=1e-6, fixSize=False, w=0, step=1, noCenter=False, x=None, abeta=None,
              bal=False, listA=None, centre_stride=1, literal=False):
    """Returns the list of output lines (header first); unevaluated rows are None."""
    pos, gen, cnt, tot, minCount = read_input(infile, nofreq, MAF, nosub, minCount, phys, Rrate)
    stat = 'B1' if nofreq else (('B0maf' if nosub else 'B2maf') if MAF else ('B0' if nosub else 'B2'))
    probs, logProbs, props = onp.neutral_per_site(spectfile, nofreq, MAF, nosub, cnt, tot)
    xs, al, As = build_grids(x, abeta, bal, listA)
    sel = onp.per_site_tables(cnt, tot, xs, al, stat, minCount)
    A_order, x_order, a_order = list(set(As)), list(set(xs)), list(set(al))
    pairs = [(xx, aa) for xx in x_order for aa in a_order]
    selmat = np.stack([sel[p] * props for p in pairs])
    if fixSize and not phys:
        phys = True                                        # v1:619-621 (message only)
    rows = onp.scan_rows(pos, gen, Rrate, fixSize=fixSize, r=w, s=step, noCenter=noCenter)
    lines = [HEADER]
    for j, row in enumerate(rows):
        if j % centre_stride:
            lines.append(None)
            continue
        if row['gap']:
            lines.append(onp.format_row(row, 0, 0, 0, 0, 0))
            continue
        if literal:
            T, xh, ah, Ah, ns = onp.calc_baller(row['lo'], row['hi'], row['t'], gen, probs, logProbs,
                                                props, sel, A_order, x_order, a_order)
        else:
            T, iA, ixa, ns = onp.calc_baller_fast(row['lo'], row['hi'], row['t'], gen, probs,
                                                  logProbs, props, selmat, A_order)
            if iA < 0:
                T, xh, ah, Ah, ns = 0., 0., 0., 0., 0.
            else:
                T = np.float64(T); xh, ah = pairs[ixa]; Ah = A_order[iA]
        lines.append(onp.format_row(row, T, xh, ah, Ah, ns))
    return lines
