"""Neutral model -- host side of SURVEY.md §8 row a3.

Mirrors ``NeutralSFS`` of the reference (/root/reference/BalLeRMix+_v1.py:180-304):
constructor ``NeutralSFS(spectfile, nofreq, MAF, nosub)``; attributes ``spect``
(dict (k, n) -> g), ``sampProps`` (dict n -> sum of g over that n),
``sampSizes``; ``get_neut_probs(data)`` fills per-site ``probs``, ``logProbs``
and ``propSizes``.  Per-site arrays are produced by a class lookup
(unique (k, n) pairs) rather than ``np.vectorize`` over sites; the values are
the same dictionary entries.

Semantics kept:
  * spect file: ``k<TAB>n<TAB>fraction``, no header; fractions are summed in file
    order into the checksum and into ``sampProps[n]``            (v1:186-214)
  * --MAF: a line with k >= n/2 + 1 is folded into (n-k, n)      (v1:195-203)
  * --MAF --noSub with k == 0, or --noSub (DAF) with k == n: message + exit(0)
                                                                  (v1:192-194,205-207)
  * the checksum must be np.isclose to 1                          (v1:218)
  * config file (B1): ``n<TAB>sub<TAB>poly``; the dict is rebuilt on every line so
    only the LAST line's n survives; checksum == 1.0 exactly     (v1:227-250)
  * every (k, n) of the data must be a key of the spectrum       (v1:281-287)
"""
import sys

import numpy as np

_MSG_SUM = 'Fraction of sites do not add up to 1! Sum = {}. Please double-check your inputs.'


def site_classes(count, total):
    """Unique (k, n) classes of the data.

    Returns (class_k int64[C], class_n int64[C], cls int32[N]) with classes sorted
    by (n, k) and ``cls[i]`` the class of site i.
    """
    count = np.asarray(count, dtype=np.int64)
    total = np.asarray(total, dtype=np.int64)
    if count.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int32)
    kmax = int(count.max()) + 1
    if count.min() < 0 or total.min() < 0:
        raise ValueError('negative allele count or sample size in input')
    key = total * kmax + count
    uniq, inv = np.unique(key, return_inverse=True)
    return uniq % kmax, uniq // kmax, inv.astype(np.int32)


class NeutralSFS:

    def __init__(self, spectfile, nofreq, MAF, nosub):
        self.spect = {}
        self.sampSizes = set()
        self.sampProps = {}
        self.probs = []
        self.logProbs = []
        self.propSizes = []
        if nofreq:
            self.readConfig(spectfile)
        else:
            self.readSpect(spectfile, MAF, nosub)

    @classmethod
    def from_spect(cls, spect):
        """Build from an in-memory {(k, n): fraction} spectrum (bench / tests; no file)."""
        self = cls.__new__(cls)
        self.spect = dict(spect)
        self.sampProps = {}
        for (k, n), f in self.spect.items():
            self.sampProps[n] = self.sampProps.get(n, 0.) + f
        self.sampSizes = set(self.sampProps)
        self.probs, self.logProbs, self.propSizes = [], [], []
        return self

    def readSpect(self, spectfile, MAF, nosub):
        g = {}
        sizes = []
        checksum = 0.
        with open(spectfile, 'r') as fh:
            for line in fh:
                f = line.strip().split('\t')
                k = int(f[0]); n = int(f[1]); frac = float(f[2])
                if MAF:
                    if nosub and k == 0:
                        print('You have chosen to compute B_0maf. Please do not account for sites '
                              'with zero counts (x==0) in your input.')
                        sys.exit()
                    if k < (n / 2 + 1):
                        g[(k, n)] = frac
                    else:
                        print('You have indicated to use minor allele frequencies (--MAF) but '
                              'provided SFS based on polarized allele frequency. This SFS will be '
                              'folded accordingly.')
                        if (n - k, n) in g:
                            g[(n - k, n)] += frac
                        else:
                            g[(n - k, n)] = frac
                else:
                    if nosub and k == n:
                        print('You have chosen to compute B_2maf. Please do not account for '
                              'substitutions (derived allele count x == n) in your input.')
                        sys.exit()
                    g[(k, n)] = frac
                checksum += frac
                sizes.append(n)
                if n not in self.sampProps:
                    self.sampProps[n] = 0.
                self.sampProps[n] += frac
        if not np.isclose(checksum, 1.):
            print(_MSG_SUM.format(checksum))
            sys.exit()
        self.spect = g
        self.sampSizes = set(sizes)

    def readConfig(self, spectfile):
        sizes = []
        checksum = 0.
        g = {}
        with open(spectfile, 'r') as fh:
            for line in fh:
                f = line.strip().split('\t')
                n = int(f[0]); s = float(f[1]); p = float(f[2])
                print('Substitutions: %s ; polymorphisms: %s' % (s, p))
                checksum += (s + p)
                g = {(0, n): s, (1, n): p}          # rebuilt per line, as v1:236
                sizes.append(n)
                if n not in self.sampProps:
                    self.sampProps[n] = 0
                self.sampProps[n] += (s + p)
        if not checksum == 1.:
            print(_MSG_SUM.format(checksum))
            sys.exit()
        self.spect = g
        self.sampSizes = set(sizes)

    def class_tables(self, class_k, class_n):
        """g(k, n) and the sample-size proportion for each class: (G[C], P[C])."""
        pairs = list(zip(class_k.tolist(), class_n.tolist()))
        if not set(pairs).issubset(self.spect.keys()):
            print('Input data includes sample counts and sizes not included in the helper file. '
                  'Please double-check your inputs.')
            sys.exit()
        G = np.array([self.spect[p] for p in pairs], dtype=np.float64)
        P = np.array([self.sampProps[p[1]] for p in pairs], dtype=np.float64)
        return G, P

    def get_neut_probs(self, InputData):
        class_k, class_n, cls = site_classes(InputData.count, InputData.total)
        G, P = self.class_tables(class_k, class_n)
        self.class_k, self.class_n, self.cls = class_k, class_n, cls
        self.classG, self.classP = G, P
        self.probs = G[cls]
        assert len(self.probs) == InputData.numSites
        with np.errstate(divide='ignore'):
            self.logProbs = np.log(self.probs)
        self.propSizes = P[cls]
