#!/usr/bin/env python3
"""Drive the C-ABI multi-GPU entry (include/blmx_mgpu.h, libblmx_mgpu.so) the way a non-Python host would:
one thread per GPU, an NCCL communicator made with ncclCommInitAll, every rank calls blmx_scan_sharded with the
same centre list, rank 0 receives the rows.  The result is compared with a single-GPU blmx_scan of the same
centres.  No torch in this process (its bundled NCCL must not meet the system libnccl the library links).

    python tools/mgpu_abi_demo.py [--gpus 2] [--sites 40000] [--centres 600]
"""
import argparse
import ctypes as C
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['BLMX_LIB'] = os.path.join(ROOT, 'ballermixplus_b200', 'libblmx_mgpu.so')   # the superset library
from ballermixplus_b200 import native                                                   # noqa: E402


def make_problem(n_sites, seed=5):
    rng = np.random.default_rng(seed)
    n_cls, n_x, n_a = 40, 5, 21
    g = np.sort(rng.random(n_sites)) * 0.02
    w = rng.random(n_cls) ** 3 + 1e-3
    w[0] += 3.0
    cls = rng.choice(n_cls, size=n_sites, p=w / w.sum()).astype(np.int32)
    G = rng.random(n_cls) * 0.1 + 1e-3
    SP = G[None, :] * 10.0 ** rng.normal(0, 0.3, (n_x * n_a, n_cls))
    A = np.array([500., 1000., 2000., 5000., 2e4, 1e5])
    return native.ScanProblem(g, cls, G, SP, A, n_x, n_a)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=2)
    ap.add_argument('--sites', type=int, default=40000)
    ap.add_argument('--centres', type=int, default=600)
    opt = ap.parse_args()
    L = native.lib()
    L.blmx_scan_sharded.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(native.Result)]
    L.blmx_shard_ranges.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]
    if native.device_count() < opt.gpus:
        print(f'SKIP: {native.device_count()} GPU(s) visible, {opt.gpus} wanted')
        return 0
    nccl = C.CDLL('libnccl.so.2')
    world = opt.gpus
    comms = (C.c_void_p * world)()
    devs = (C.c_int * world)(*range(world))
    rc = nccl.ncclCommInitAll(comms, world, devs)
    assert rc == 0, f'ncclCommInitAll -> {rc}'

    prob = make_problem(opt.sites)
    n = len(prob.genpos)
    rng = np.random.default_rng(9)
    c = np.sort(rng.choice(n, size=opt.centres, replace=False))
    t = np.ascontiguousarray(prob.genpos[c])
    lo = np.zeros(len(c), np.int64)
    hi = np.full(len(c), n - 1, np.int64)
    hi[::9] = c[::9] + 300                                      # some index-cut windows

    with native.Scanner(device=0).load(prob) as sc:
        want = sc.scan(t, lo, hi)
        begin, end = np.zeros(world, np.int64), np.zeros(world, np.int64)
        rc = L.blmx_shard_ranges(sc._h, world, len(c), native._ptr(t), native._ptr(lo), native._ptr(hi),
                                 native._ptr(begin), native._ptr(end))
        assert rc == 0, L.blmx_last_error()
    assert begin[0] == 0 and end[-1] == len(c) and np.all(begin[1:] == end[:-1]), (begin, end)

    out = [np.zeros(len(c)), *(np.zeros(len(c), np.int32) for _ in range(4))]
    res = native.Result(*(native._ptr(a) for a in out))
    errors = []

    def rank_main(r):
        try:
            with native.Scanner(device=r).load(prob) as sc:
                rc = L.blmx_scan_sharded(sc._h, r, world, comms[r], len(c), native._ptr(t), native._ptr(lo),
                                         native._ptr(hi), C.byref(res) if r == 0 else None)
                if rc != 0:
                    errors.append((r, rc, L.blmx_last_error()))
        except Exception as exc:                                 # noqa: BLE001
            errors.append((r, repr(exc)))

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for r in range(world):
        nccl.ncclCommDestroy(comms[r])
    assert not errors, errors
    for a, b in zip(out, want):
        assert np.array_equal(a, b), 'rows of the sharded scan differ from the single-GPU scan'
    print(f'OK: {len(c)} centres on {world} GPUs (ranges {list(zip(begin.tolist(), end.tolist()))}), '
          f'{int(np.sum(want[1] >= 0))} rows with T > 0, identical to the single-GPU scan')
    return 0


if __name__ == '__main__':
    sys.exit(main())
