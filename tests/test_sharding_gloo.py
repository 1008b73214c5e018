"""Multi-GPU host logic on CPU: cost-balanced centre shards + the single gather, world_size 2
and 3 over gloo (127.0.0.1), with the C oracle standing in for the per-rank CUDA scan."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from ballermixplus_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _problem():
    from ballermixplus_b200.problem import build_problem
    argv, _ = util.scan_cases()['ex1_B2_w15_s7p5']
    opt, data, neutral, grid, sel = util.host_objects(argv)
    prob, order = build_problem(data, neutral, sel, grid)
    n = data.numSites
    c = np.arange(0, n, 3)
    lo = np.maximum(0, c - 40)
    hi = np.minimum(n - 1, c + 25)
    return prob, data.genPos[c], lo, hi


def _oracle_scan_fn(prob):
    from oracle import oracle_c

    def fn(t, lo, hi):
        T, iA, ixa, ns, _ = oracle_c.scan(prob.genpos, prob.cls, prob.G, prob.SP, prob.A, t, lo, hi, n_threads=1)
        ix = np.where(ixa >= 0, ixa // prob.n_a, -1).astype(np.int32)
        ia = np.where(ixa >= 0, ixa % prob.n_a, -1).astype(np.int32)
        return (torch.from_numpy(T), torch.from_numpy(iA), torch.from_numpy(ix), torch.from_numpy(ia),
                torch.from_numpy(ns))
    return fn


def _worker(rank, world, port, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        prob, t, lo, hi = _problem()
        got = sharding.scan_sharded(prob.genpos, prob.A, t, lo, hi, rank, world, _oracle_scan_fn(prob), dist, torch)
        if rank == 0:
            torch.save([g.clone() for g in got], out_path)
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_scan_equals_single_process(world, tmp_path):
    out = str(tmp_path / 'rows.pt')
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    prob, t, lo, hi = _problem()
    ref = _oracle_scan_fn(prob)(t, lo, hi)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)


def test_partition_is_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 1000, size=5000).astype(float)
    costs[:500] = 1.0                       # cheap centres at a sequence end
    for world in (1, 2, 4, 8):
        parts = sharding.partition(costs, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(costs)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        sums = np.array([costs[b:e].sum() for b, e in parts])
        assert sums.max() <= costs.sum() / world + costs.max()
    assert sharding.partition(np.ones(3), 8)[-1][1] == 3        # more ranks than centres
    assert sharding.partition(np.zeros(0), 2) == [(0, 0), (0, 0)]


def test_pack_unpack_roundtrip():
    T = torch.tensor([0.0, 1.5, -np.inf, 3.3027966623106693], dtype=torch.float64)
    idx = [torch.tensor(v, dtype=torch.int32) for v in ([-1, 2, 3, 4], [-1, 0, 9, 1], [-1, 50, 0, 7], [0, 756, 3, 16])]
    back = sharding.unpack_rows(sharding.pack_rows(T, *idx, torch), torch)
    assert torch.equal(back[0], T)
    for a, b in zip(back[1:], idx):
        assert torch.equal(a, b)


def test_centre_costs_follow_alpha_reach():
    g = np.arange(1000) * 1e-5
    t = g[[0, 500, 999]]
    lo, hi = np.zeros(3, np.int64), np.full(3, 999, np.int64)
    c = sharding.centre_costs(g, t, lo, hi, [1e4, 1e5])
    assert c[1] > c[0] and c[1] > c[2]      # interior centres see both sides
    assert abs(c[0] - c[2]) <= 2
