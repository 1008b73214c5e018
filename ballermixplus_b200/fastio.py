"""ctypes binding of include/blmx_io.h: C++ fast paths for reading the input table and writing
the output rows.  Pure host code; every function returns None when the
library is missing or declines the input, and the callers then run the Python code that mirrors
the reference line for line -- so results never depend on which path ran."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libblmx_io.so')
IO_SYMBOLS = ('blmx_io_last_error', 'blmx_io_count_rows', 'blmx_io_read_sites', 'blmx_io_write_rows',
              'blmx_io_format_doubles')
_lib = None


def lib():
    global _lib
    if _lib is None and os.path.exists(LIB_PATH) and not os.environ.get('BLMX_NO_FASTIO'):
        L = C.CDLL(LIB_PATH)
        L.blmx_io_last_error.restype = C.c_char_p
        L.blmx_io_count_rows.argtypes = [C.c_char_p, C.POINTER(C.c_int64)]
        L.blmx_io_read_sites.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        L.blmx_io_write_rows.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.c_char_p), C.c_int32,
                                         C.POINTER(C.c_char_p), C.c_int32]
        L.blmx_io_format_doubles.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def read_sites(path, use_phys, rrate, strict_columns=False):
    """-> (position, genPos, count, total) or None if the fast path is unavailable / declines."""
    L = lib()
    if L is None:
        return None
    n = C.c_int64(0)
    if L.blmx_io_count_rows(os.fsencode(path), C.byref(n)) != 0:
        return None
    pos = np.empty(n.value, np.int64)
    gen = np.empty(n.value, np.float64)
    cnt = np.empty(n.value, np.int64)
    tot = np.empty(n.value, np.int64)
    rc = L.blmx_io_read_sites(os.fsencode(path), n.value, int(bool(use_phys)), float(rrate),
                              int(bool(strict_columns)), _p(pos), _p(gen),
                              _p(cnt), _p(tot))
    if rc != 0:
        return None
    return pos, gen, cnt, tot


def write_rows(path, header, physpos, genpos, T, iA, ix, ia, nsites, A_text, x_text, a_text):
    """Write the output file; returns False if the fast path is unavailable."""
    L = lib()
    if L is None:
        return False
    physpos = np.ascontiguousarray(physpos, np.int64)
    genpos = np.ascontiguousarray(genpos, np.float64)
    T = np.ascontiguousarray(T, np.float64)
    iA, ix, ia, nsites = (np.ascontiguousarray(v, np.int32) for v in (iA, ix, ia, nsites))

    def texts(items):
        arr = (C.c_char_p * max(1, len(items)))()
        for i, s in enumerate(items):
            arr[i] = s.encode()
        return arr

    At, xt, at = texts(A_text), texts(x_text), texts(a_text)
    rc = L.blmx_io_write_rows(os.fsencode(path), header.encode(), len(T), _p(physpos), _p(genpos), _p(T), _p(iA),
                              _p(ix), _p(ia), _p(nsites), At, len(A_text), xt, len(x_text), at, len(a_text))
    if rc != 0:
        raise OSError(L.blmx_io_last_error().decode())
    return True


def format_doubles(values):
    """str(np.float64(v)) for each v via the C++ formatter (used by the tests)."""
    L = lib()
    if L is None:
        return None
    v = np.ascontiguousarray(values, np.float64)
    cap = 40 * len(v) + 16
    buf = C.create_string_buffer(cap)
    n = C.c_int64(0)
    if L.blmx_io_format_doubles(_p(v), len(v), buf, cap, C.byref(n)) != 0:
        raise ValueError(L.blmx_io_last_error().decode())
    return buf.raw[:n.value].decode().split('\n')[:-1]
