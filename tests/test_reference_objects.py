"""Only where /root/reference exists (the build container): the product against the reference's
OWN objects -- per-class tables vs the reference's per-site tables, and the calcBaller seam fed
with reference objects."""
import importlib.util
import os
import sys

import numpy as np
import pytest

import util

REF = '/root/reference/BalLeRMix+_v1.py'
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason='reference tree not present (GPU box)')


@pytest.fixture(scope='module')
def ref():
    spec = importlib.util.spec_from_file_location('ballermix_ref', REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ref_objects(ref, argv):
    from ballermixplus_b200.cli import build_parser
    opt = build_parser().parse_args(util.abs_paths(argv))
    with util.quiet():
        data = ref.InputData(opt.infile, opt.nofreq, opt.MAF, opt.nosub, opt.minCount, phys=opt.phys, Rrate=opt.Rrate)
        neutral = ref.NeutralSFS(opt.spectfile, opt.nofreq, opt.MAF, opt.nosub)
        neutral.get_neut_probs(data)
        grid = ref.Grids(opt.x, opt.abeta, opt.bal, opt.pos, opt.seqA, opt.listA)
        sel = ref.NormalizedBetaBinom(data, grid, opt.nofreq, opt.MAF, opt.nosub)
    return data, neutral, grid, sel


@pytest.mark.parametrize('name', ['Example1_B2', 'Example2_B2maf', 'Example1_B1', 'ex2_B0_s5',
                                  'Example2_B0maf_1kb-2site', 'synth_mixed_n_B2_s20'])
def test_host_objects_equal_reference_objects(ref, name):
    argv, _ = util.scan_cases()[name]
    rdata, rneutral, rgrid, rsel = _ref_objects(ref, argv)
    opt, data, neutral, grid, sel = util.host_objects(argv)
    assert np.array_equal(data.position, rdata.position) and np.array_equal(data.genPos, rdata.genPos)
    assert np.array_equal(data.count, rdata.count) and np.array_equal(data.total, rdata.total)
    assert data.minCount == rdata.minCount and data.sampSizes == rdata.sampSizes
    assert np.array_equal(neutral.probs, rneutral.probs) and np.array_equal(neutral.propSizes, rneutral.propSizes)
    assert grid.x == rgrid.x and grid.A == rgrid.A and grid.abeta == rgrid.abeta
    assert [type(v) for v in grid.abeta] == [type(v) for v in rgrid.abeta]
    for key, per_site in rsel.normProbs.items():
        assert np.array_equal(sel.get(*key), per_site), key        # bit-identical, all 510 tables


def test_problem_from_reference_objects_is_identical(ref):
    """build_problem accepts the reference's objects (what the calcBaller drop-in relies on)."""
    from ballermixplus_b200.problem import build_problem
    argv, _ = util.scan_cases()['Example1_B2maf']
    rdata, rneutral, rgrid, rsel = _ref_objects(ref, argv)
    opt, data, neutral, grid, sel = util.host_objects(argv)
    a, oa = build_problem(rdata, rneutral, rsel, rgrid)
    b, ob = build_problem(data, neutral, sel, grid)
    for f in ('genpos', 'cls', 'G', 'SP', 'A'):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert (oa.A, oa.x, oa.a) == (ob.A, ob.x, ob.a)


def test_reference_calcballer_vs_oracle_on_strided_centres(ref):
    """The reference function itself against the numpy oracle (extra pin beside the goldens)."""
    from oracle import oracle_np
    argv, _ = util.scan_cases()['ex1_B2_w15_s7p5']
    rdata, rneutral, rgrid, rsel = _ref_objects(ref, argv)
    A, x, a = list(set(rgrid.A)), list(set(rgrid.x)), list(set(rgrid.abeta))
    for i in (0, 101, 377, 756):
        lo, hi = max(0, i - 15), min(rdata.numSites - 1, i + 16)
        with util.quiet():
            want = ref.calcBaller(np.arange(lo, hi + 1), rdata.genPos[i], rdata, rneutral, rsel, rgrid)
        got = oracle_np.calc_baller(lo, hi, rdata.genPos[i], rdata.genPos, rneutral.probs, rneutral.logProbs,
                                    rneutral.propSizes, rsel.normProbs, A, x, a)
        assert want[1:] == got[1:]
        assert abs(want[0] - got[0]) <= 1e-12 * max(1., abs(want[0]))
