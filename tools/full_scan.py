#!/usr/bin/env python3
"""Full-size runs of the two synthetic targets of BASELINE.json -- every site a centre, not a sample:

  cfg4  one chromosome, 500 k informative sites (35 Mb), n = 200, --usePhysPos --rec 1e-8, A = --rangeA
        1000,10900,100 x default x / alpha grids (51 000 grid points): run through the drop-in CLI
        (input + spect files on disk, `python -m ballermixplus_b200 ...`, output file), one GPU.
  cfg5  whole genome, 10 M sites / 22 chromosomes, same grids: every rank (one per GPU, torchrun) scans its
        cost-balanced contiguous share of the 10 M centres through the C ABI, one NCCL gather to rank 0.

    python tools/full_scan.py --config cfg4
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/full_scan.py --config cfg5

Rank 0 then checks >= 64 randomly chosen centres against the CPU oracle (CLR within 1e-9, identical argmax and
nSites) and writes profiles/r2_full_<config>.json: wall time of the scan, centres, centres x grid points / s.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                     # noqa: E402


def oracle_check(prob, t, lo, hi, rows, picks, threads):
    """rows = (T, iA, ix, ia, ns) arrays of ALL centres; picks = indices to verify."""
    from oracle import oracle_c
    rT, rA, rxa, rn, _ = oracle_c.scan(prob.genpos, prob.cls, prob.G, prob.SP, prob.A, t[picks], lo[picks], hi[picks],
                                       n_threads=threads)
    T, iA, ix, ia, ns = (np.asarray(a)[picks] for a in rows)
    xa = np.where(iA >= 0, ix * prob.n_a + ia, -1)
    rel = np.abs(T - rT) / np.maximum(np.abs(rT), 1.)
    return {'centres': int(len(picks)), 'rows_T_gt_0': int(np.sum(rA >= 0)), 'max_rel_dT': float(rel.max()),
            'argmax_mismatch': int(np.sum((iA != rA) | (xa != rxa))), 'nsites_mismatch': int(np.sum(ns != rn))}


def run_cfg4(opt):
    """The CLI end to end on a 500 k-site file."""
    chrom = bench.make_chromosome(opt.sites or 500_000, seed=12345)
    work = tempfile.mkdtemp(prefix='blmx_cfg4_')
    prefix = os.path.join(work, 'chr')
    t0 = time.perf_counter()
    bench.write_reference_inputs(prefix, chrom)
    t_write = time.perf_counter() - t0
    out = prefix + '_out.txt'
    cmd = [sys.executable, '-m', 'ballermixplus_b200', '-i', prefix + '.txt', '--spect', prefix + '_spect.txt',
           '-o', out, '--usePhysPos', '--rec', str(bench.REC_RATE), '--rangeA', bench.RANGE_A]
    t0 = time.perf_counter()
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if res.returncode != 0:
        sys.exit(res.stdout[-2000:] + res.stderr[-2000:])
    stamps = [ln for ln in res.stdout.splitlines() if 'Start computing' in ln or 'Scan finished' in ln]
    # rows of the output file against the oracle on random centres
    prob = bench.make_problem([chrom])[0]
    n = len(prob.genpos)
    with open(out) as fh:
        lines = fh.read().splitlines()
    assert len(lines) == n + 1, (len(lines), n)
    from ballermixplus_b200 import Grids
    from ballermixplus_b200.problem import GridOrder
    order = GridOrder(Grids(None, None, False, False, bench.RANGE_A, None))
    text = {'A': [f'{v}' for v in order.A], 'x': [f'{v}' for v in order.x], 'a': [f'{v}' for v in order.a]}
    rng = np.random.default_rng(7)
    picks = np.sort(rng.choice(n, size=opt.check, replace=False))
    T = np.zeros(n); iA = np.full(n, -1, np.int64); ix = iA.copy(); ia = iA.copy(); ns = np.zeros(n, np.int64)
    for j in picks:
        f = lines[j + 1].split('\t')
        T[j] = float(f[2])
        if not (f[3] == '0.0' and f[4] == '0.0' and f[5] == '0.0'):
            ix[j], ia[j], iA[j], ns[j] = text['x'].index(f[3]), text['a'].index(f[4]), text['A'].index(f[5]), int(f[6])
    t = prob.genpos
    lo, hi = np.zeros(n, np.int64), np.full(n, n - 1, np.int64)
    parity = oracle_check(prob, t, lo, hi, (T, iA, ix, ia, ns), picks, bench.host_threads())
    n_grid = prob.n_x * prob.n_a * len(prob.A)
    return {'config': 'cfg4: synthetic chr22-scale B2 scan, %d sites, n=200, --usePhysPos --rec 1e-8, --rangeA %s, '
                      'every site a centre, through the CLI (files in, file out)' % (n, bench.RANGE_A),
            'command': ' '.join(cmd[1:]), 'n_gpus': 1, 'centres': n, 'grid_points_per_centre': n_grid,
            'wall_s_cli_process': wall, 'cli_stamps': stamps, 'input_write_s': t_write,
            'value_over_process_wall': n * n_grid / wall, 'unit': 'centre*gridpoint/s', 'parity': parity}


def run_cfg5(opt, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ballermixplus_b200 import native, sharding
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    t0 = time.perf_counter()
    problems = bench.make_problem(bench.make_genome(opt.sites or 10_000_000))
    shard = bench.Shard(problems, opt.stride, rank, world, torch, dev, sharding)
    scanners = {c: native.Scanner(device=local_rank).load(problems[c]) for c, _ in shard.mine}
    t_setup = time.perf_counter() - t0
    stream = torch.cuda.current_stream()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    o = 0
    for c, _ in shard.mine:
        t, lo, hi = shard.d_in[c]
        n = t.shape[0]
        scanners[c].scan_device(n, t.data_ptr(), lo.data_ptr(), hi.data_ptr(), shard.d_T[o:].data_ptr(),
                                *[x[o:].data_ptr() for x in shard.d_idx], stream=stream.cuda_stream)
        o += n
    rows = sharding.pack_rows(shard.d_T, *shard.d_idx, torch)
    got = sharding.gather_rows(rows, shard.counts, rank, world, dist, torch)
    host = got.cpu() if got is not None else None
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    wall = time.perf_counter() - w0
    if world > 1:
        tmax = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms, wall = (float(v) for v in tmax.tolist())
    out = None
    if rank == 0:
        T, iA, ix, ia, ns = (a.numpy() for a in sharding.unpack_rows(host, torch))
        n_grid = problems[0].n_x * problems[0].n_a * len(problems[0].A)
        rng = np.random.default_rng(11)
        picks = np.sort(rng.choice(shard.total, size=opt.check, replace=False))
        par = {'centres': 0, 'rows_T_gt_0': 0, 'max_rel_dT': 0.0, 'argmax_mismatch': 0, 'nsites_mismatch': 0}
        for c in range(len(problems)):
            b, e = shard.offs[c], shard.offs[c + 1]
            mine = picks[(picks >= b) & (picks < e)]
            if not len(mine):
                continue
            t, lo, hi = shard.plans[c]
            r = oracle_check(problems[c], t, lo, hi, tuple(a[b:e] for a in (T, iA, ix, ia, ns)), mine - b,
                             bench.host_threads())
            for k in par:
                par[k] = max(par[k], r[k]) if k == 'max_rel_dT' else par[k] + r[k]
        out = {'config': 'cfg5: synthetic whole-genome B2 scan, %d sites / %d chromosomes, n=200, --usePhysPos --rec '
                         '1e-8, --rangeA %s, centre stride %d, through the C ABI (device buffers) + one NCCL gather'
                         % (sum(len(p.genpos) for p in problems), len(problems), bench.RANGE_A, opt.stride),
               'n_gpus': world, 'centres': shard.total, 'grid_points_per_centre': n_grid,
               'scan_s_device_events_max_over_ranks': ms * 1e-3, 'scan_s_wall_max_over_ranks': wall,
               'setup_s_rank0 (synthetic data, tables, load of the problems)': t_setup,
               'value': shard.total * n_grid / (ms * 1e-3), 'unit': 'centre*gridpoint/s',
               'rows_T_gt_0_all_centres': int(np.sum(iA >= 0)), 'parity': par}
    for s in scanners.values():
        s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', choices=['cfg4', 'cfg5'], required=True)
    ap.add_argument('--sites', type=int, default=0, help='override the number of sites (smoke tests)')
    ap.add_argument('--stride', type=int, default=1, help='cfg5: centre stride (1 = every site)')
    ap.add_argument('--check', type=int, default=64, help='centres verified against the CPU oracle')
    ap.add_argument('--out', default=None)
    opt = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    res = run_cfg4(opt) if opt.config == 'cfg4' else run_cfg5(opt, rank, world, local_rank)
    if rank == 0 and res is not None:
        ok = (res['parity']['max_rel_dT'] <= 1e-9 and res['parity']['argmax_mismatch'] == 0
              and res['parity']['nsites_mismatch'] == 0)
        res['parity']['ok'] = bool(ok)
        path = opt.out or os.path.join(ROOT, 'gpurun_out', f'r2_full_{opt.config}.json')
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, 'w') as fh:
            json.dump(res, fh, indent=1)
        print(json.dumps(res))
        if not ok:
            sys.exit('full_scan: PARITY FAILED')


if __name__ == '__main__':
    main()
