"""CPU ORACLE (test infrastructure, not product code) -- numpy restatement.

A plain-numpy restatement of the reference's algorithm for the hot path
(/root/reference/BalLeRMix+_v1.py, "v1" below).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the
product path (ballermixplus_b200/) never does and fails loudly without its CUDA
library.

PARITY IS PINNED: tests/test_oracle_golden.py checks this module against the
reference's seven shipped golden scans (tests/golden/ref_out, copied from
/root/reference/test/output) and against outputs of the unmodified reference
run in the build container for the flags no shipped golden covers
(tests/golden/gen, produced by tests/golden/make_golden.py).

Functions and the reference lines they follow
  per_site_tables      v1:319-433  NormalizedBetaBinom (per SITE, literally, as an
                                   independent check of the product's per-class tables)
  neutral_per_site     v1:183-304  NeutralSFS.readSpect/readConfig/get_neut_probs
  calc_baller          v1:436-507  calcBaller (one centre; sweep A, x, a; strict argmax)
  scan_rows            v1:510-640  the four Scan window modes -> (t, lo, hi) per row
  format_row           v1:540,574,591,607  output row formatting
"""
import math

import numpy as np
from scipy.stats import betabinom


# ----------------------------------------------------------------------------- a3
def neutral_per_site(spectfile, nofreq, MAF, nosub, count, total):
    """-> (probs[N], logProbs[N], propSizes[N]); follows v1:183-304."""
    spect, samp = {}, {}
    with open(spectfile) as fh:
        for line in fh:
            f = line.strip().split('\t')
            if nofreq:                                   # readConfig v1:227-250
                n = int(f[0]); s = float(f[1]); p = float(f[2])
                spect = {(0, n): s, (1, n): p}           # re-created every line (v1:236)
                samp[n] = samp.get(n, 0) + (s + p)
            else:                                        # readSpect v1:183-223
                k = int(f[0]); n = int(f[1]); fr = float(f[2])
                if MAF and not (k < n / 2 + 1):
                    spect[(n - k, n)] = spect[(n - k, n)] + fr if (n - k, n) in spect else fr
                else:
                    spect[(k, n)] = fr
                samp[n] = samp.get(n, 0.) + fr
    probs = np.array([spect[(int(k), int(n))] for k, n in zip(count, total)], dtype=np.float64)
    props = np.array([samp[int(n)] for n in total], dtype=np.float64)
    with np.errstate(divide='ignore'):
        return probs, np.log(probs), props


# ----------------------------------------------------------------------------- a4
def per_site_tables(count, total, xs, alphas, stat, minCount):
    """dict (x, a) -> float64[N], evaluated per site as v1:319-359 does."""
    count = np.asarray(count); total = np.asarray(total)
    out = {}

    def bb(j, n, x, a):
        return betabinom(n, a, a / x - a).pmf(j)        # v1:316,366-371

    def raw(k, n, x, a):                                 # v1:375-396
        if stat == 'B1':
            return np.where(k == 0, bb(n, n, x, a), (1. - bb(n, n, x, a) - bb(n, n, x, a)))
        if stat in ('B2', 'B0'):
            return bb(k, n, x, a)
        p = bb(k, n, x, a) + bb(n - k, n, x, a)
        if n % 2 == 0:
            p = np.where(k == int(n / 2), p / 2, p)
        return p

    def excluded(n):                                     # v1:399-433
        m = minCount
        if stat in ('B2', 'B1'):
            return np.arange(m)
        if stat == 'B2maf':
            return np.concatenate((np.arange(m), np.arange(n - m + 1, n)))
        if stat == 'B0':
            return np.concatenate((np.arange(m), np.array([n])))
        return np.concatenate((np.arange(m), np.arange(n - m + 1, n + 1)))

    for x in xs:
        for a in alphas:
            full = np.zeros(len(count))
            for n in set(total.tolist()):
                idx = np.where(total == n)
                k = count[idx]
                folded = 0.5 * (raw(k, n, x, a) + raw(k, n, 1. - x, a))
                e = excluded(n)
                base = 1. - np.sum(0.5 * (bb(e, n, x, a) + bb(e, n, 1. - x, a)))
                full[idx] = folded / base
            out[(x, a)] = full
    return out


# ----------------------------------------------------------------------------- a5
def calc_baller(lo, hi, t, genPos, probs, logProbs, propSizes, sel, A_order, x_order, a_order):
    """One centre, literally v1:436-507.

    ``sel[(x, a)]`` are per-site normalised selection probabilities, the window is the
    inclusive index range [lo, hi], grids are visited in the given order (the
    reference's ``set`` order) with a strict ``>`` update from Tmax = 0.
    Returns [T, x, a, A, nSites] (all 0. when no grid point has T > 0).
    """
    dist = np.abs(genPos - t)                                           # v1:446
    best = [0., 0., 0., 0., 0.]
    n = len(genPos)
    in_window = np.zeros(n, dtype=bool)
    in_window[lo:hi + 1] = True
    for A in A_order:                                                   # v1:453
        alphas = np.exp(-A * dist)                                      # v1:454
        sub = np.flatnonzero((alphas >= 1e-8) & (genPos != t) & in_window)   # v1:455-457
        if len(sub) == 0:
            continue
        al = alphas[sub]
        neut = probs[sub]
        cl_neut = np.sum(logProbs[sub])                                 # v1:497
        prop = propSizes[sub]
        for x in x_order:
            for a in a_order:
                s = sel[(x, a)][sub] * prop                             # v1:479,492
                mix = al * s + (1. - al) * neut                         # v1:494
                with np.errstate(divide='ignore', invalid='ignore'):
                    T = 2 * (np.sum(np.log(mix)) - cl_neut)             # v1:496-499
                if T > best[0]:                                         # v1:501
                    best = [T, x, a, A, len(sub)]
    return best


def calc_baller_fast(lo, hi, t, genPos, probs, logProbs, propSizes, selmat, A_order):
    """Same arithmetic as calc_baller with the (x, a) loops vectorised.

    ``selmat`` is float64[n_xa, N] = sel[(x, a)] * propSizes in visiting order.
    Returns (T, iA, ixa, nSites) with iA = ixa = -1 when nothing has T > 0.
    """
    dist = np.abs(genPos - t)
    bT, bA, bxa, bn = 0., -1, -1, 0
    for iA, A in enumerate(A_order):
        alphas = np.exp(-A * dist)
        ok = (alphas >= 1e-8) & (genPos != t)
        sub = lo + np.flatnonzero(ok[lo:hi + 1])
        if len(sub) == 0:
            continue
        al = alphas[sub]
        with np.errstate(divide='ignore', invalid='ignore'):
            mix = al[None, :] * selmat[:, sub] + ((1. - al) * probs[sub])[None, :]
            T = 2 * (np.sum(np.log(mix), axis=1) - np.sum(logProbs[sub]))
        Tc = np.where(np.isnan(T), -np.inf, T)
        j = int(np.argmax(Tc))                     # first maximum == strict '>' in visiting order
        if Tc[j] > bT:
            bT, bA, bxa, bn = float(Tc[j]), iA, j, len(sub)
    return bT, bA, bxa, bn


# ----------------------------------------------------------------------------- a6
def scan_rows(position, genPos, Rrate, fixSize=False, r=0, s=1, noCenter=False):
    """Rows of the scan: list of dicts with the centre, the inclusive window and the
    two leading output fields, for each of the four modes of v1:513-640."""
    N = len(position)
    rows = []
    if fixSize and noCenter:                                            # v1:513-545
        w = float(r)
        start = int(math.floor(2 * float(position[0]) / w) * (w / 2))
        end = start + s; midpos = start + s / 2
        si = ei = 0
        while midpos <= position[-1]:
            while position[si] < start:
                si += 1
            while (ei + 1) < N:
                if position[ei] < end:
                    ei += 1
                else:
                    break
            if si >= ei:
                rows.append(dict(gap=True, mid=midpos, t=midpos * Rrate))
            else:
                rows.append(dict(gap=False, t=midpos * Rrate, lo=si, hi=ei,
                                 f0=f'{midpos}', f1=f'{midpos * Rrate}'))
            start += s; midpos += s; end += s
    elif fixSize:                                                       # v1:549-577
        w = float(r)
        i = 0; si = 0; ei = 0
        while i < N:
            ts = position[i]
            start = max(0, ts - w / 2); end = min(ts + w / 2, position[-1])
            while position[si] < start:
                si += 1
            while ei < N:
                if position[ei] < end:
                    ei += 1
                else:
                    break
            ei = min(ei, N - 1)
            rows.append(dict(gap=False, t=genPos[i], lo=si, hi=ei, f0=f'{ts}', f1=f'{genPos[i]}'))
            i += int(s)
    elif r != 0:                                                        # v1:580-594
        i = 0
        while i < N:
            w = np.arange(max(0, i - r), min(N - 1, i + r + 1) + 1, dtype=int)
            rows.append(dict(gap=False, t=genPos[int(i)], lo=int(w[0]), hi=int(w[-1]),
                             f0=f'{position[int(i)]}', f1=f'{genPos[int(i)]}'))
            i += s
    else:                                                               # v1:598-610
        i = 0
        while i < N:
            rows.append(dict(gap=False, t=genPos[int(i)], lo=0, hi=N - 1,
                             f0=f'{position[int(i)]}', f1=f'{genPos[int(i)]}'))
            i += int(s)
    return rows


def format_row(row, T, x, a, A, nsites):
    """v1:535 for gap rows, v1:540/574/591/607 otherwise (same f-string fields)."""
    if row['gap']:
        return '%d\t%g\t0\tNA\tNA\tNA\t0\n' % (row['mid'], row['t'])
    return f"{row['f0']}\t{row['f1']}\t{T}\t{x}\t{a}\t{A}\t{nsites}\n"
