#!/usr/bin/env python3
"""Benchmark of the CLR scan: centres x grid-points / s on synthetic whole-genome data.

Workload (BASELINE.json configs[4]): 10 M informative sites over 22
chromosomes (sizes proportional to the human autosomes, mean spacing 70 nt), n = 200 at
every site, P(substitution) = 0.7 else P(k) ~ 1/k, genPos = pos * 1e-8 (--usePhysPos --rec
1e-8), B2 statistic, default x (10) and alpha (51) grids, A = --rangeA 1000,10900,100
(100 values) -> 51 000 grid points per centre, default window mode (alpha >= 1e-8).
Scanning every site of that genome takes minutes even at roofline, so a "step" scans the
whole genome with the CLI's own --step flag: every S-th informative site is a centre
(S = 1024/N for N GPUs: per-GPU work is fixed, "weak" scaling).  Every centre still sees
the full 10 M-site data, exactly like `-s S` in the reference (v1:603-608).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo, CUDA
    python bench.py --impl reference [...]                        # CPU arm (oracle port, all threads)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HUMAN_AUTOSOME_MB = [249.25, 243.20, 198.02, 191.15, 180.92, 171.12, 159.14, 146.36, 141.21, 135.53,
                     135.01, 133.85, 115.17, 107.35, 102.53, 90.35, 81.20, 78.08, 59.13, 63.03, 48.13, 51.30]
N_SAMPLE = 200
SPACING = 70
REC_RATE = 1e-8
RANGE_A = '1000,10900,100'
BASE_STRIDE = 1024
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12        # 37.2: 148 SM x 64 DFMA/clk x 1.965 GHz
# algorithmic FP64 flops (FMA = 2, MUL/ADD = 1; libdevice exp = 31, log = 47)
FLOP_SINGLE = 3.0          # (1-alpha) + alpha*R (FMA) and the running product (MUL), per site and grid point
FLOP_GROUPED = 2.25        # four sites: 4 FMA (quartic in R) + 1 MUL
FLOP_EXP = 34.0            # alpha = exp(-A*|g - t|): once per (centre, A, site) evaluated site by site
FLOP_QUAD_COEF = 23.0 / 4  # f0..f4 of a quartic from its four alphas (one lane), per grouped site
FLOP_LOG = 47.0 + 3.0      # one log + exponent fold-in per (centre, A, grid point)
FLOP_BLOCK = 32 * (FLOP_EXP + 2.0)   # far field: one exp + one FMA on each of the 32 moment lanes, per block visit
FLOP_EDGE = 9.0            # far field: five moments of a block-remainder site (besides its exp)
FLOP_POLY = 2.0            # far field: one Horner FMA per polynomial term and grid point


def algorithmic_flops(cnt, n_xa, n_items):
    """FP64 flops the formulation needs for the work the kernel's own counters report:
    sites evaluated per grid point (singly or four at a time), one exp per site that is looked at
    individually, the far-field block visits / remainder sites / polynomial terms, one log per grid point."""
    direct = cnt['pairs'] - cnt['far_sites']
    grouped = direct - cnt['single']
    return (n_xa * (FLOP_SINGLE * cnt['single'] + FLOP_GROUPED * grouped) + FLOP_QUAD_COEF * grouped
            + FLOP_EXP * (direct + cnt['edge_sites']) + FLOP_EDGE * cnt['edge_sites']
            + FLOP_BLOCK * cnt['far_blocks'] + FLOP_POLY * n_xa * cnt['far_terms'] + FLOP_LOG * n_xa * n_items)


# ------------------------------------------------------------------------ synthetic data
def make_chromosome(n_sites, seed, n=N_SAMPLE, spacing=SPACING):
    """positions = sorted sample without replacement from [1, L)."""
    rng = np.random.default_rng(seed)
    length = max(n_sites * spacing, n_sites + 2)
    pos = np.sort(rng.choice(length - 1, size=n_sites, replace=False)).astype(np.int64) + 1
    is_sub = rng.random(n_sites) < 0.7
    w = 1. / np.arange(1, n)
    k = np.where(is_sub, n, rng.choice(np.arange(1, n), size=n_sites, p=w / w.sum())).astype(np.int64)
    return {'pos': pos, 'k': k, 'n': n}


def genome_sizes(total_sites):
    w = np.array(HUMAN_AUTOSOME_MB) / sum(HUMAN_AUTOSOME_MB)
    sizes = np.floor(w * total_sites).astype(np.int64)
    sizes[0] += total_sites - sizes.sum()
    return sizes.tolist()


def make_genome(total_sites, seed=12345):
    return [make_chromosome(int(s), seed + c) for c, s in enumerate(genome_sizes(total_sites))]


def make_problem(chroms, range_a=RANGE_A):
    """Host precompute with the product's own classes (what the CLI does, without files):
    spectrum = empirical class frequencies of the concatenation (what --getSpect writes)."""
    from ballermixplus_b200 import Grids, InputData, NeutralSFS, NormalizedBetaBinom
    from ballermixplus_b200.native import ScanProblem
    from ballermixplus_b200.problem import GridOrder
    n = chroms[0]['n']
    counts = np.zeros(n + 1, np.int64)
    for c in chroms:
        counts += np.bincount(c['k'], minlength=n + 1)
    total = counts.sum()
    ks = np.flatnonzero(counts)
    spect = {(int(k), n): float(counts[k]) / float(total) for k in ks}
    classes = InputData.from_arrays(np.arange(len(ks)), np.arange(len(ks)) * 1e-6, ks, np.full(len(ks), n))
    neutral = NeutralSFS.from_spect(spect)
    grid = Grids(None, None, False, False, range_a, None)
    sel = NormalizedBetaBinom(classes, grid, False, False, False)
    G, P = neutral.class_tables(sel.class_k, sel.class_n)
    order = GridOrder(grid)
    SP = np.stack([sel.classProbs[(x, a)] * P for x in order.x for a in order.a])
    A = np.array([float(v) for v in order.A])
    class_of_k = np.full(n + 1, -1, np.int64)
    class_of_k[sel.class_k] = np.arange(len(sel.class_k))
    out = []
    for c in chroms:
        out.append(ScanProblem(c['pos'] * REC_RATE, class_of_k[c['k']].astype(np.int32), G, SP, A,
                               len(order.x), len(order.a)))
    return out


def plan_centres(problems, stride):
    """Default mode, `-s stride`: (chromosome, t, lo, hi) for every stride-th site (v1:598-610)."""
    plans = []
    for p in problems:
        n = len(p.genpos)
        idx = np.arange(0, n, stride)
        plans.append((p.genpos[idx].copy(), np.zeros(len(idx), np.int64), np.full(len(idx), n - 1, np.int64)))
    return plans


# ------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.device)],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        with open(self.path) as fh:
            for line in fh:
                f = [v.strip() for v in line.split(',')]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(names, f[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'power_w_max': float(max(pw)),
                'samples': len(sm), 'reasons': sorted(reasons)}


def cpu_sample_size(problems, plans, threads, target_s=15.0):
    """Centres for about `target_s` seconds of CPU work on this host: time two centres first
    (an interior centre of the 10 M-site workload is 6.6e8 log evaluations, parallelised over
    (centre, A) tasks), then scale."""
    secs, centres, _ = run_cpu_oracle(problems, plans, cpu_sample(problems, plans, 2, threads), threads)
    per_centre = max(secs / max(centres, 1), 1e-3)
    return int(min(4096, max(2, round(target_s / per_centre))))


def cpu_sample(problems, plans, n_centres, threads):
    """A bounded sample for the CPU arm: centres spread evenly over the whole genome."""
    sizes = np.array([len(p[0]) for p in plans])
    picks = np.linspace(0, sizes.sum() - 1, n_centres).astype(np.int64)
    bounds = np.concatenate(([0], np.cumsum(sizes)))
    sample = []
    for c in range(len(plans)):
        mine = picks[(picks >= bounds[c]) & (picks < bounds[c + 1])] - bounds[c]
        if len(mine):
            sample.append((c, mine))
    return sample


def run_cpu_oracle(problems, plans, sample, threads):
    """Literal calcBaller (oracle/oracle_c.c, OpenMP) on the sample -> (seconds, centres, site pairs)."""
    from oracle import oracle_c
    t0 = time.perf_counter()
    centres = pairs = 0
    for c, idx in sample:
        p = problems[c]
        t, lo, hi = plans[c]
        res = oracle_c.scan(p.genpos, p.cls, p.G, p.SP, p.A, t[idx], lo[idx], hi[idx], n_threads=threads)
        centres += len(idx)
        pairs += res[4]
    return time.perf_counter() - t0, centres, pairs


def run_numpy_oracle_one_centre(problems, plans):
    """The vectorised numpy restatement (oracle/oracle_np.py, what the reference's numpy/scipy path
    costs once its Python-list indexing overhead is removed), ONE interior centre, one thread."""
    from oracle import oracle_np
    c = int(np.argmax([len(p.genpos) for p in problems]))
    p = problems[c]
    t = plans[c][0]
    tj = float(t[len(t) // 2])
    reach = 18.420680743952367 / float(p.A.min()) * 1.001
    lo = int(np.searchsorted(p.genpos, tj - reach, 'left'))
    hi = int(np.searchsorted(p.genpos, tj + reach, 'right'))
    g = p.genpos[lo:hi]
    cls = p.cls[lo:hi]
    probs = p.G[cls]
    selmat = np.ascontiguousarray(p.SP[:, cls])
    t0 = time.perf_counter()
    oracle_np.calc_baller_fast(0, len(g) - 1, tj, g, probs, np.log(probs), np.ones_like(probs), selmat, p.A)
    secs = time.perf_counter() - t0
    n_grid = p.n_x * p.n_a * len(p.A)
    return {'value': n_grid / secs, 'unit': 'centre*gridpoint/s', 'cores': 1, 'kind': 'port',
            'sample': f'1 interior centre, {n_grid} grid points, numpy restatement (oracle/oracle_np.py), {secs:.1f} s'}


# ---------------------------------------------------------------------------------- arms
def reference_arm(opt, rank):
    """CPU arm: the oracle's C port of calcBaller on all host threads (the reference itself is a
    Python script that cannot travel to the GPU box; oracle/ is its restatement)."""
    if rank != 0:
        return
    from oracle import oracle_c
    threads = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    threads = min(threads, oracle_c.max_threads()) if oracle_c.max_threads() > 0 else threads
    chroms = make_genome(opt.sites)
    problems = make_problem(chroms)
    stride = max(1, BASE_STRIDE // opt.gpus)
    plans = plan_centres(problems, stride)
    n_grid = problems[0].n_x * problems[0].n_a * len(problems[0].A)
    per_step = opt.cpu_centres or cpu_sample_size(problems, plans, threads, target_s=12.0)
    sample = cpu_sample(problems, plans, per_step, threads)
    for _ in range(opt.warmup):
        run_cpu_oracle(problems, plans, cpu_sample(problems, plans, 1, threads), threads)
    secs = centres = pairs = 0
    for _ in range(opt.steps):
        s, c, p = run_cpu_oracle(problems, plans, sample, threads)
        secs += s; centres += c; pairs += p
    value = centres * n_grid / secs
    line = {
        'impl': 'reference', 'metric': 'centres x grid-points / s (B2 scan)', 'value': value,
        'unit': 'centre*gridpoint/s', 'n_gpus': opt.gpus, 'steps': opt.steps, 'warmup': opt.warmup,
        'ms_per_step': secs / opt.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(opt, stride, problems, plans),
        'cpu_baseline': {'value': value, 'unit': 'centre*gridpoint/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{per_step} centres per step spread evenly over the genome, all '
                                   f'{n_grid} grid points each, literal calcBaller in C (oracle/oracle_c.c), OpenMP',
                         'site_evals_per_s': pairs * problems[0].n_x * problems[0].n_a / secs},
        'e2e': {'value': value, 'unit': 'centre*gridpoint/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(opt, stride, problems, plans):
    n_sites = int(sum(len(p.genpos) for p in problems))
    return {
        'workload': f'cfg5 synthetic whole-genome B2 scan: {n_sites} sites / {len(problems)} chromosomes, n=200, '
                    f'--usePhysPos --rec 1e-8, --rangeA {RANGE_A} x default x/alpha grids, default window '
                    f'mode, centres = every {stride}-th site (-s {stride})',
        'sites': n_sites, 'chromosomes': len(problems), 'centres_per_step': int(sum(len(p[0]) for p in plans)),
        'grid_points_per_centre': problems[0].n_x * problems[0].n_a * len(problems[0].A),
        'centre_stride': stride, 'parallelism': f'centre-range shards x{opt.gpus}',
        'l2': 'site arrays (200 MB at 10 M sites) exceed the 126 MB L2; an extra 256 MB write flushes L2 between steps',
    }


def cuda_arm(opt, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ballermixplus_b200 import native, sharding

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    chroms = make_genome(opt.sites)
    problems = make_problem(chroms)
    del chroms
    stride = max(1, BASE_STRIDE // world)
    plans = plan_centres(problems, stride)
    n_xa = problems[0].n_x * problems[0].n_a
    n_A = len(problems[0].A)
    n_grid = n_xa * n_A

    # cost-balanced contiguous shard of the global centre list
    costs = np.concatenate([sharding.centre_costs(p.genpos, pl[0], pl[1], pl[2], p.A)
                            for p, pl in zip(problems, plans)])
    parts = sharding.partition(costs, world)
    counts = [e - b for b, e in parts]
    begin, end = parts[rank]
    offs = np.concatenate(([0], np.cumsum([len(pl[0]) for pl in plans])))
    mine = []        # (chromosome, slice into its plan)
    for c in range(len(plans)):
        b, e = max(begin, offs[c]), min(end, offs[c + 1])
        if e > b:
            mine.append((c, slice(int(b - offs[c]), int(e - offs[c]))))
    n_mine = end - begin
    total_centres = int(offs[-1])

    # resident problems + device buffers
    scanners = {}
    for c, _ in mine:
        scanners[c] = native.Scanner(device=local_rank, group=opt.group, farfield=opt.farfield).load(problems[c])
        scanners[c].set_option('timing', 1)
    stream = torch.cuda.current_stream().cuda_stream
    host_in = {c: tuple(np.ascontiguousarray(a[sl]) for a in plans[c]) for c, sl in mine}
    pinned = {c: tuple(torch.from_numpy(a).pin_memory() for a in host_in[c]) for c in host_in}
    d_in = {c: tuple(x.to(dev) for x in pinned[c]) for c in pinned}
    d_T = torch.zeros(n_mine, dtype=torch.float64, device=dev)
    d_idx = [torch.zeros(n_mine, dtype=torch.int32, device=dev) for _ in range(4)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def scan_resident():
        o = 0
        for c, _ in mine:
            t, lo, hi = d_in[c]
            n = t.shape[0]
            scanners[c].scan_device(n, t.data_ptr(), lo.data_ptr(), hi.data_ptr(), d_T[o:].data_ptr(),
                                    *[x[o:].data_ptr() for x in d_idx], stream=stream)
            o += n
        rows = sharding.pack_rows(d_T, *d_idx, torch)
        return sharding.gather_rows(rows, counts, rank, world, dist, torch)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------
    for _ in range(opt.warmup):
        scan_resident()
        flush.zero_()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def collect():
        """Per-launch kernel events and work counters of the LAST step, summed over this rank's scanners."""
        tot = {'pairs': 0, 'single': 0, 'far_blocks': 0, 'far_terms': 0, 'far_sites': 0, 'edge_sites': 0,
               'quads': 0, 'launches': 0}
        k_ms, k_n = 0.0, 0
        for c, _ in mine:
            kms, kn = scanners[c].kernel_ms()
            k_ms += kms; k_n += kn
            for key, v in scanners[c].counters_all().items():
                tot[key] = tot.get(key, 0) + v
        return tot, k_ms, k_n

    sync_all()
    e0.record()
    for _ in range(opt.steps):
        gathered = scan_resident()
        flush.zero_()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    cnt, kernel_ms, kernel_launches = collect()
    pairs, single, lib_launches = cnt['pairs'], cnt['single'], cnt['launches']
    clocks = sampler.stop() if rank == 0 else None

    # the same workload with every site evaluated directly (option farfield = 0): the FP64-bound kernel
    direct = None
    if opt.farfield and world == 1:
        for c, _ in mine:
            scanners[c].set_option('farfield', 0)
        scan_resident()
        sync_all()
        e0.record()
        scan_resident()
        e1.record()
        sync_all()
        d_ms = e0.elapsed_time(e1)
        d_cnt, d_kms, d_kn = collect()
        direct = (d_ms, d_cnt, d_kms, d_kn)
        for c, _ in mine:
            scanners[c].set_option('farfield', 1)
    if world > 1:
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
        agg = torch.tensor([pairs, single, lib_launches, kernel_ms, kernel_launches], dtype=torch.float64, device=dev)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        pairs_all, single_all, launches_all = int(agg[0].item()), int(agg[1].item()), int(agg[2].item())
    else:
        pairs_all, single_all, launches_all = pairs, single, lib_launches
    value = total_centres * n_grid * opt.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ---------------------------------
    def scan_e2e():
        o = 0
        for c, _ in mine:
            scanners[c].load(problems[c])                       # H2D of sites + tables, class sort
            t, lo, hi = (x.to(dev, non_blocking=True) for x in pinned[c])   # H2D of the centres
            n = t.shape[0]
            scanners[c].scan_device(n, t.data_ptr(), lo.data_ptr(), hi.data_ptr(), d_T[o:].data_ptr(),
                                    *[x[o:].data_ptr() for x in d_idx], stream=stream)
            o += n
        rows = sharding.pack_rows(d_T, *d_idx, torch)
        g = sharding.gather_rows(rows, counts, rank, world, dist, torch)
        return g.cpu() if g is not None else None                 # D2H of every row on rank 0

    scan_e2e()
    sync_all()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(opt.steps):
        host_rows = scan_e2e()
    e1.record()
    sync_all()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    if world > 1:
        tmax = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = float(tmax.item())
    h2d = sum(problems[c].h2d_bytes + 24 * (sl.stop - sl.start) for c, sl in mine)
    d2h = 24 * total_centres if rank == 0 else 0
    if world > 1:
        agg = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        h2d, d2h = int(agg[0].item()), int(agg[1].item())
    e2e_value = total_centres * n_grid * opt.steps / (e2e_ms * 1e-3)

    if rank == 0:
        # sanity: the gathered rows are complete and decoded
        T, iA, ix, ia, ns = sharding.unpack_rows(host_rows, torch)
        assert T.shape[0] == total_centres and bool((iA >= -1).all()) and bool((ns >= 0).all())

        peak_tf, peak_mhz = native.measure_fp64_peak(local_rank, 0.5)
        n_items = n_mine * n_A
        flops = algorithmic_flops(cnt, n_xa, n_items)
        k_s = kernel_ms * 1e-3
        achieved = flops / k_s / 1e12 if k_s > 0 else None
        table_bytes = 8.0 * n_xa * problems[0].G.shape[0] * n_items          # D rows a warp may touch
        site_bytes = 8.0 * pairs
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get('dram_bytes_per_launch')
        hbm_peak = None
        mp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(mp):
            with open(mp) as fh:
                hbm_peak = json.load(fh).get('hbm_gbs')
        line = {
            'metric': 'centres x grid-points / s (B2 scan)', 'value': value, 'unit': 'centre*gridpoint/s',
            'n_gpus': world, 'steps': opt.steps, 'warmup': opt.warmup, 'ms_per_step': ms / opt.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(opt, stride, problems, plans),
            'site_evals_per_s': pairs_all * n_xa / (ms / opt.steps * 1e-3),
            'mean_sites_per_centre_A': pairs_all / max(1, total_centres * n_A),
            'e2e': {'value': e2e_value, 'unit': 'centre*gridpoint/s', 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_ms / opt.steps},
            'gpu_launches': int(launches_all * opt.steps),
            'clocks': clocks,
            'roofline': {
                'bound': 'fp64', 'kernel': 'scan_kernel<16,%d,%s>' % (opt.group, 'far' if opt.farfield else 'direct'),
                'achieved': achieved,
                'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': (achieved / peak_tf) if achieved and peak_tf else None,
                'peak_source': 'measured live: register-resident DFMA loop (blmx_measure_fp64_peak); '
                               'MEASURED_PEAKS.json has no FP64 entry',
                'peak_nominal': FP64_NOMINAL_TFLOPS,
                'frac_of_nominal': (achieved / FP64_NOMINAL_TFLOPS) if achieved else None,
                'kernel_ms_per_launch': kernel_ms / max(1, kernel_launches), 'launches_timed': kernel_launches,
                'algorithmic_flops_per_launch': flops / max(1, kernel_launches),
                'sites_single_frac': single / max(1, pairs), 'sites_far_frac': cnt['far_sites'] / max(1, pairs),
                'traffic': traffic,
                'hbm': {'algorithmic_bytes_per_launch': (site_bytes + table_bytes) / max(1, kernel_launches),
                        'achieved_gbs': (site_bytes + table_bytes) / k_s / 1e9 if k_s > 0 else None,
                        'peak_gbs': hbm_peak,
                        'note': 'upper bound (every class row counted); the path is FP64-bound'},
            },
        }
        line['mode'] = ('farfield: far sites through power sums of alpha (DESIGN.md §3.4)' if opt.farfield
                        else 'direct: every site evaluated per grid point')
        if direct is not None:
            d_ms, d_cnt, d_kms, d_kn = direct
            d_flops = algorithmic_flops(d_cnt, n_xa, n_items)
            d_ach = d_flops / (d_kms * 1e-3) / 1e12
            line['direct'] = {
                'value': total_centres * n_grid / (d_ms * 1e-3), 'unit': 'centre*gridpoint/s', 'ms_per_step': d_ms,
                'roofline': {'bound': 'fp64', 'kernel': 'scan_kernel<16,%d,direct>' % opt.group, 'achieved': d_ach,
                             'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': d_ach / peak_tf,
                             'kernel_ms_per_launch': d_kms / max(1, d_kn), 'launches_timed': d_kn,
                             'algorithmic_flops_per_launch': d_flops / max(1, d_kn)},
                'note': 'same step with option farfield = 0 (one warm-up + one timed step)'}
        if world == 1 and not opt.no_cpu:
            threads = len(os.sched_getaffinity(0))
            n_cpu = opt.cpu_centres or cpu_sample_size(problems, plans, threads)
            sample = cpu_sample(problems, plans, n_cpu, threads)
            secs, centres, cpairs = run_cpu_oracle(problems, plans, sample, threads)
            line['cpu_baseline'] = {
                'value': centres * n_grid / secs, 'unit': 'centre*gridpoint/s', 'cores': threads, 'kind': 'port',
                'sample': f'{centres} centres spread evenly over the genome, all {n_grid} grid points each, '
                          f'literal calcBaller in C (oracle/oracle_c.c), OpenMP on {threads} threads, {secs:.1f} s',
                'site_evals_per_s': cpairs * n_xa / secs}
            line['cpu_baseline_numpy'] = run_numpy_oracle_one_centre(problems, plans)
        print(json.dumps(line), flush=True)
    for s in scanners.values():
        s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--sites', type=int, default=10_000_000, help='informative sites in the synthetic genome')
    ap.add_argument('--group', type=int, default=4, choices=[1, 4])
    ap.add_argument('--farfield', type=int, default=1, choices=[0, 1],
                    help='1 (default): far sites enter through power sums; 0: every site evaluated directly')
    ap.add_argument('--cpu-centres', type=int, default=0, help='centres in the CPU sample (default: one per thread)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    opt = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if opt.impl == 'reference':
        reference_arm(opt, rank)
        return
    if world != opt.gpus and world > 1:
        opt.gpus = world
    cuda_arm(opt, rank, world, local_rank)


if __name__ == '__main__':
    main()
