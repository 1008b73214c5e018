"""Input reader -- host side of SURVEY.md §8 row a1.

Mirrors the public surface of the reference's ``InputData`` class
(/root/reference/BalLeRMix+_v1.py:8-131): same constructor arguments, same
attributes (``position``, ``genPos``, ``count``, ``total``, ``numSites``,
``minCount``, ``Rrate``, ``sampSizes``), same variant rules.  The implementation
is column-wise numpy instead of per-line list appends.

Semantics kept:
  * the first line is always a header and is skipped          (v1:83,116)
  * physPos = int(float(col0)), k = int(col2), n = int(col3)  (v1:88,121)
  * genPos  = float(col0)*Rrate with --usePhysPos, else float(col1) (v1:103,124)
  * B1 (--noFreq): the first row whose k is not 0/1 switches on "translate"
    mode; that row and every later row get k = (k != n)       (v1:91-100)
  * --noSub: rows with k == n are dropped                     (v1:41-50)
  * --MAF: k > n/2 is folded to n-k; minCount = min(k > 0)    (v1:53-59)
  * otherwise any k == 0 prints the reference's message and exits with
    status 0; minCount = min(k)                               (v1:60-74)
  * B1 keeps the constructor's ``minCount`` (the CLI value)   (v1:16)
"""
import sys

import numpy as np

from . import fastio

_MSG_TRANSLATE = ('Input includes different variant counts despite choosing not to use allele '
                  'frequencies (with --noFreq). All sites with counts smaller than substitutions '
                  'will be considered as polymorphic. All sites with identical counts as sample '
                  'sizes will be substitutions.')
_MSG_ZERO = ('Please make sure to include derived allele frequency only. Sites with zero derived '
             'alleles (x==0) should not be included in your input.')


def read_site_table(infile, use_phys, Rrate):
    """Parse the 4-column tab-separated input (header skipped).

    Returns (position int64[N], genPos float64[N], k int64[N], n int64[N]).
    """
    fast = fastio.read_sites(infile, use_phys, Rrate)      # C++ reader; None -> the loop below
    if fast is not None:
        return fast
    pos, gen, kk, nn = [], [], [], []
    col = 0 if use_phys else 1
    with open(infile, 'r') as fh:
        next(fh, None)
        for line in fh:
            f = line.strip().split('\t')
            pos.append(int(float(f[0])))
            gen.append(float(f[col]))
            kk.append(int(f[2]))
            nn.append(int(f[3]))
    position = np.array(pos, dtype=np.int64) if pos else np.zeros(0, np.int64)
    g = np.array(gen, dtype=np.float64) if gen else np.zeros(0, np.float64)
    if use_phys:
        # v1:103/124 evaluates float(c0)*1*Rrate + float(c0)*0 ; the +0.0 is exact
        g = g * 1 * Rrate + g * 0
    else:
        # float(c1)*0*Rrate + float(c1)*1
        g = g * 0 * Rrate + g * 1
    count = np.array(kk, dtype=np.int64) if kk else np.zeros(0, np.int64)
    total = np.array(nn, dtype=np.int64) if nn else np.zeros(0, np.int64)
    return position, g, count, total


class InputData:
    """Same constructor and attributes as the reference class (v1:8-76)."""

    def __init__(self, infile, nofreq=False, MAF=False, nosub=False, minCount=1, phys=False,
                 Rrate=1e-6):
        self.minCount = minCount
        self.Rrate = Rrate
        position, genPos, count, total = read_site_table(infile, bool(phys), Rrate)
        self.numSites = int(len(count))

        if nofreq:
            # B1: polymorphism -> 1, substitution -> 0, once a non-binary count is seen
            nonbinary = np.flatnonzero((count != 0) & (count != 1))
            if nonbinary.size:
                print(_MSG_TRANSLATE)
                first = int(nonbinary[0])
                count = count.copy()
                count[first:] = (count[first:] != total[first:]).astype(np.int64)
            # the reference leaves --minCount as given on this path (a str from argparse
            # crashes it, v1:724 -> v1:401); we accept anything int() accepts.
            self.minCount = int(minCount)
        else:
            stat = '%s%s' % (['B_2', 'B_0'][bool(nosub)], ['', 'MAF'][bool(MAF)])
            if nosub and np.any(count == total):
                print(f'You have chosen to compute {stat}. Substitutions (x==n) in the input '
                      f'will be ignored.')
                keep = count != total
                position, genPos, count, total = position[keep], genPos[keep], count[keep], total[keep]
                self.numSites = int(len(count))
            if MAF:
                over = count > total / 2
                if np.any(over):
                    print(f'Input data includes non-MAF site/s (frequency >= 0.5) despite choosing '
                          f'to use {stat} (with --MAF). These frequencies will be folded for '
                          f'following analyses.')
                    count = np.where(over, total - count, count)
                self.minCount = count[count > 0].min()
            else:
                if np.any(count == 0):
                    print(_MSG_ZERO)
                    sys.exit()
                self.minCount = count.min()

        self.position = position
        self.genPos = genPos
        self.count = count
        self.total = total
        self.sampSizes = set(total.tolist())

    @classmethod
    def from_arrays(cls, position, genPos, count, total, minCount=None, Rrate=1e-6):
        """Build an instance from already-filtered arrays (used by bench / tests)."""
        self = cls.__new__(cls)
        self.position = np.ascontiguousarray(position, dtype=np.int64)
        self.genPos = np.ascontiguousarray(genPos, dtype=np.float64)
        self.count = np.ascontiguousarray(count, dtype=np.int64)
        self.total = np.ascontiguousarray(total, dtype=np.int64)
        self.numSites = int(len(self.count))
        self.Rrate = Rrate
        self.minCount = self.count.min() if minCount is None else minCount
        self.sampSizes = set(self.total.tolist())
        return self
