"""Search grids -- host side of SURVEY.md §8 row a2.

Mirrors ``Grids`` of the reference (/root/reference/BalLeRMix+_v1.py:134-175):
same constructor ``Grids(x, abeta, bal, pos, seqA, listA)`` and the same three
attributes ``x``, ``A``, ``abeta`` holding Python lists whose element *types*
(int vs float) are kept, because they are visible in the output file
(SURVEY.md appendix A.6).

The scan iterates ``set(grid)`` (v1:453,473,474), so duplicates vanish and exact
ties resolve in CPython set-iteration order; ``scan_order()`` returns the three
lists in exactly that order.

Two flags have no reference behaviour because the reference raises on them
(SURVEY.md appendix A.2).  Semantics chosen here, and documented in DESIGN.md:
  * ``--rangeA min,max,step``: v1:169-171 calls ``range(float)`` and names an
    undefined variable.  Intended meaning: ``min + step*i`` for
    ``i = 0 .. floor((max-min)/step)`` (floats), which is what this builds.
  * ``--findPos``: v1:155 puts x = 1.0 on the grid, so ``1-x = 0`` divides by
    zero at v1:316.  Here grid values with x == 0 or x == 1 are dropped.
"""
import math


def _default_abeta():
    return ([0.001, 0.01, 0.05, 0.1, 0.2, 0.5, 0.8] + [i for i in range(1, 10)]
            + [5 * i for i in range(1, 20)] + [10 * i for i in range(10, 21)]
            + [300, 500, 1e3, 1e4, 1e6, 1e9])


def _default_A():
    return ([100 * i for i in range(1, 12)] + [200 * i for i in range(6, 13)]
            + [500 * i for i in range(5, 10)] + [1000 * i for i in range(5, 11)] + [1e6, 1e8])


class Grids:

    def __init__(self, x, abeta, bal, pos, seqA, listA):
        if x is not None:
            xs = [float(x)]
        else:
            xs = [.05 * i for i in range(1, 11)]

        if abeta is not None:
            try:
                alphas = [float(abeta)]
            except (TypeError, ValueError):
                print(f'The value for "a" provided ({abeta}) is not legitimate. Using the default '
                      f'grid instead.')
                alphas = _default_abeta()
        elif bal:
            alphas = _default_abeta()[7:]          # v1:152: the default grid without the a < 1 values
        elif pos:
            alphas = _default_abeta()[:7]          # v1:154
            xs = [v for v in (.1 * i for i in range(1, 11)) if 0. < v < 1.]
        else:
            alphas = _default_abeta()

        if not seqA and not listA:
            As = _default_A()
        elif listA:
            As = [float(v) for v in listA.split(',')]
        else:
            Amin, Amax, Astep = [float(v) for v in seqA.split(',')]
            if not Astep > 0 or Amax < Amin:
                raise ValueError(f'--rangeA expects <Amin>,<Amax>,<Astep> with Astep > 0 and '
                                 f'Amax >= Amin; got {seqA!r}')
            npts = int(math.floor((Amax - Amin) / Astep + 1e-9)) + 1
            As = [Amin + Astep * i for i in range(npts)]

        self.x = xs
        self.A = As
        self.abeta = alphas

    def scan_order(self):
        """(A, x, abeta) as lists in the order the reference's loops visit them."""
        return list(set(self.A)), list(set(self.x)), list(set(self.abeta))
