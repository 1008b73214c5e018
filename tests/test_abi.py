"""The C-ABI library loads and exports every symbol include/blmx.h declares (no GPU needed),
and the product fails loudly -- no CPU fallback -- when there is no CUDA device."""
import os
import re

import numpy as np
import pytest

import util
from ballermixplus_b200 import native


def _declared():
    with open(os.path.join(util.ROOT, 'include', 'blmx.h')) as fh:
        text = re.sub(r'/\*.*?\*/', '', fh.read(), flags=re.S)
    return sorted(set(re.findall(r'\b(blmx_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 12
    L = native.lib()
    for n in names:
        assert hasattr(L, n), n
    assert sorted(native.ABI_SYMBOLS) == names
    assert L.blmx_abi_version() == 2


def test_mgpu_library_exports_its_header():
    """libblmx_mgpu.so (built where nccl.h exists) = every symbol of blmx.h plus the two of blmx_mgpu.h."""
    import ctypes
    path = os.path.join(util.ROOT, 'ballermixplus_b200', 'libblmx_mgpu.so')
    if not os.path.exists(path):
        pytest.skip('libblmx_mgpu.so not built (no nccl.h)')
    with open(os.path.join(util.ROOT, 'include', 'blmx_mgpu.h')) as fh:
        text = re.sub(r'/\*.*?\*/', '', fh.read(), flags=re.S)
    extra = sorted(set(re.findall(r'\b(blmx_[a-z0-9_]+)\s*\(', text)))
    assert extra == ['blmx_scan_sharded', 'blmx_shard_ranges']
    try:
        L = ctypes.CDLL(path)
    except OSError as exc:
        pytest.skip(f'cannot load libblmx_mgpu.so here: {exc}')
    for n in extra + _declared():
        assert hasattr(L, n), n


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(util.ROOT, 'ballermixplus_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.h', '.cpp')):
                with open(os.path.join(dirpath, f)) as fh:
                    text = fh.read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'liboracle' not in text, f


def test_no_gpu_means_loud_failure():
    try:
        have = native.device_count() > 0
    except native.BlmxError:
        have = False
    if have:
        pytest.skip('a CUDA device is present')
    with pytest.raises(native.BlmxError):
        native.Scanner(device=0)
    from ballermixplus_b200 import calcBaller
    argv, _ = util.scan_cases()['ex1_B2_fixgrid']
    opt, data, neutral, grid, sel = util.host_objects(argv)
    with pytest.raises(native.BlmxError):
        calcBaller(np.arange(10), data.genPos[3], data, neutral, sel, grid)
