"""CPU ORACLE (test infrastructure, not product code) -- ctypes wrapper of oracle_c.c.

Build with ``make -C oracle`` (done by __graft_entry__.build()).  Used by tests, smoke()
and bench.py's cpu_baseline / --impl reference legs only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, 'liboracle.so')
_lib = None


def build():
    subprocess.run(['make', '-C', _HERE], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            build()
        L = C.CDLL(_PATH)
        L.oracle_scan_floor.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.POINTER(C.c_uint64), C.c_double]
        _lib = L
    return _lib


def max_threads():
    return lib().oracle_max_threads()


def scan(genpos, cls, G, SP, A, t, lo, hi, n_threads=0, report_all=False):
    """Literal calcBaller over a batch of centres.

    SP is [n_xa, n_classes]; returns (T, iA, ixa, nsites, site_pairs) with -1 indices
    where no grid point has T > 0.  ``report_all``: start the maximum from -inf instead of the
    reference's 0 (v1:451), i.e. report the best grid point even when its T <= 0 (the product's
    diagnostic option of the same name).
    """
    genpos = np.ascontiguousarray(genpos, np.float64)
    cls = np.ascontiguousarray(cls, np.int32)
    G = np.ascontiguousarray(G, np.float64)
    SP = np.ascontiguousarray(SP, np.float64)
    A = np.ascontiguousarray(A, np.float64)
    t = np.ascontiguousarray(t, np.float64)
    lo = np.ascontiguousarray(lo, np.int64)
    hi = np.ascontiguousarray(hi, np.int64)
    n = len(t)
    T = np.zeros(n)
    iA = np.zeros(n, np.int32)
    ixa = np.zeros(n, np.int32)
    ns = np.zeros(n, np.int32)
    pairs = C.c_uint64(0)

    def p(a):
        return a.ctypes.data_as(C.c_void_p)

    rc = lib().oracle_scan_floor(len(genpos), p(genpos), p(cls), len(G), p(G), p(SP), SP.shape[0], len(A), p(A),
                                 n, p(t), p(lo), p(hi), p(T), p(iA), p(ixa), p(ns), int(n_threads),
                                 C.byref(pairs), -np.inf if report_all else 0.0)
    if rc != 0:
        raise MemoryError('oracle_scan failed')
    return T, iA, ixa, ns, pairs.value
