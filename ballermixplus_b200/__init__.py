"""B200-native BalLeRMix+ composite-likelihood-ratio scan.

Host side mirrors the reference script's objects (InputData, Grids, NeutralSFS,
NormalizedBetaBinom, calcBaller, Scan, getSpect, getConfig, main); the scan itself
runs in hand-written sm_100a CUDA kernels behind the C ABI of include/blmx.h
(ballermixplus_b200/libblmx.so).  No CPU fallback.
"""
from .grids import Grids
from .helpers import getConfig, getSpect
from .inputs import InputData
from .neutral import NeutralSFS
from .selection import NormalizedBetaBinom
from .scan import DeviceScan, Scan, calcBaller, clear_cache
from .cli import main

__all__ = ['Grids', 'InputData', 'NeutralSFS', 'NormalizedBetaBinom', 'DeviceScan', 'Scan',
           'calcBaller', 'clear_cache', 'getSpect', 'getConfig', 'main']
