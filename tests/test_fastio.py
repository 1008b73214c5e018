"""C++ text I/O fast paths against the Python code that mirrors the
reference: same arrays from the reader, same bytes from the writer.  No GPU."""
import glob
import os

import numpy as np
import pytest

import util
from ballermixplus_b200 import fastio, inputs
from ballermixplus_b200.problem import GridOrder
from ballermixplus_b200.scan import HEADER, format_rows, write_scan
from ballermixplus_b200 import Grids, InputData, windows


def python_reader(path, use_phys, rate, monkeypatch):
    monkeypatch.setattr(fastio, 'read_sites', lambda *a, **k: None)
    out = inputs.read_site_table(path, use_phys, rate)
    monkeypatch.undo()
    return out


def test_library_exports_declared_symbols():
    import re
    with open(os.path.join(util.ROOT, 'include', 'blmx_io.h')) as fh:
        text = re.sub(r'/\*.*?\*/', '', fh.read(), flags=re.S)
    names = sorted(set(re.findall(r'\b(blmx_io_[a-z0-9_]+)\s*\(', text)))
    L = fastio.lib()
    assert L is not None
    for n in names:
        assert hasattr(L, n), n
    assert sorted(fastio.IO_SYMBOLS) == names


@pytest.mark.parametrize('use_phys,rate', [(False, 1e-6), (True, 1e-6), (True, 1e-8)])
def test_reader_matches_python_on_reference_inputs(use_phys, rate, monkeypatch):
    files = sorted(glob.glob(os.path.join(util.GOLD, 'data', 'Example*.txt')) +
                   glob.glob(os.path.join(util.GOLD, 'data', 'synth_*.txt')))
    assert len(files) >= 8
    for f in files:
        fast = fastio.read_sites(f, use_phys, rate)
        assert fast is not None, f
        slow = python_reader(f, use_phys, rate, monkeypatch)
        for a, b in zip(fast, slow):
            assert a.dtype == b.dtype and np.array_equal(a, b), f


def test_reader_odd_but_valid_formatting(tmp_path, monkeypatch):
    f = tmp_path / 'odd.txt'
    f.write_bytes(b'physPos\tgenPos\tx\tn\r\n'
                  b'12.9\t1.2e-07\t50\t50\r\n'               # CRLF, decimal position (truncated)
                  b' 165\t 1.65E-06 \t+3\t50  \n'             # spaces inside fields, explicit sign
                  b'2.22e2\t.00000222\t12\t50\textra\n'        # exponent position, extra column ignored
                  b'499\t4.99e-06\t38\t50')                     # no final newline
    for use_phys in (False, True):
        fast = fastio.read_sites(str(f), use_phys, 1e-6)
        slow = python_reader(str(f), use_phys, 1e-6, monkeypatch)
        assert fast is not None
        for a, b in zip(fast, slow):
            assert np.array_equal(a, b)
    assert fast[0].tolist() == [12, 165, 222, 499]
    assert fastio.read_sites(str(f), True, 1.0, strict_columns=True) is None     # the 5-column line


@pytest.mark.parametrize('body', [b'1_000\t0.1\t1\t10\n', b'10\t0.1\t1.0\t10\n', b'10\t0.1\t1\n',
                                  b'0x10\t0.1\t1\t10\n', b'10\tabc\t1\t10\n', b'\n', b'inf\t0.1\t1\t10\n'])
def test_reader_declines_what_it_cannot_take_verbatim(tmp_path, body):
    """Anything Python's int()/float() would treat differently, or reject, is left to Python."""
    f = tmp_path / 'bad.txt'
    f.write_bytes(b'h\n' + body)
    assert fastio.read_sites(str(f), False, 1e-6) is None


def test_empty_and_header_only_files(tmp_path):
    f = tmp_path / 'e.txt'
    f.write_bytes(b'physPos\tgenPos\tx\tn\n')
    out = fastio.read_sites(str(f), False, 1e-6)
    assert out is not None and all(len(a) == 0 for a in out)


def test_double_formatting_is_numpy_str():
    rng = np.random.default_rng(4)
    vals = np.concatenate([rng.random(50000) * 10.0 ** rng.integers(-320, 308, 50000),
                           -rng.random(2000) * 10.0 ** rng.integers(-20, 20, 2000),
                           rng.integers(0, 10 ** 17, 2000).astype(float),
                           [0.0, -0.0, 1e16, 9999999999999998.0, 1e-4, 9.999e-5, 1e-5, 1.2e-07, 31.21810547602786,
                            0.005085999999999999, 1e22, 1.7976931348623157e308, 5e-324, np.inf, -np.inf, np.nan,
                            0.15000000000000002, 1000000000.0, 100000000.0, 0.1, 123456.0]])
    got = fastio.format_doubles(vals)
    assert got == [str(np.float64(v)) for v in vals]
    assert got == [repr(float(v)) for v in vals]


@pytest.mark.parametrize('flags', [dict(), dict(r=7, s=3.0), dict(fixSize=True, r=800, s=2.0)])
def test_writer_bytes_equal_python_formatter(flags, tmp_path):
    rng = np.random.default_rng(9)
    n = 900
    pos = np.sort(rng.choice(np.arange(3, 90000), n, replace=False))
    data = InputData.from_arrays(pos, pos * 1.3e-7, np.ones(n, int), np.full(n, 10), Rrate=1e-6)
    with util.quiet():
        plan = windows.make_plan(data, phys=True, **flags)
    order = GridOrder(Grids(None, None, False, False, None, None))
    m = len(plan)
    T = rng.random(m) * 10.0 ** rng.integers(-8, 4, m)
    iA = rng.integers(0, len(order.A), m).astype(np.int32)
    ix = rng.integers(0, len(order.x), m).astype(np.int32)
    ia = rng.integers(0, len(order.a), m).astype(np.int32)
    ns = rng.integers(1, 60000, m).astype(np.int32)
    lose = rng.random(m) < 0.1                      # the reference's all-zero rows
    T[lose] = 0
    iA[lose] = ix[lose] = ia[lose] = -1
    ns[lose] = 0
    out = str(tmp_path / 'fast.txt')
    write_scan(out, plan, order, T, iA, ix, ia, ns)
    want = HEADER + ''.join(format_rows(plan, order, T, iA, ix, ia, ns))
    with open(out) as fh:
        assert fh.read() == want
    assert '\t0.0\t0.0\t0.0\t0.0\t0.0\n' in want and '\t1000000000.0\t' in want
