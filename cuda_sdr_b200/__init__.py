"""B200-native (sm_100a) streaming DSP hot path of kernrj/cuda-sdr.

Python here is harness glue only (ctypes over the C-ABI in include/, torch for device memory, streams
and torch.distributed).  The product is the native code in csrc/ (kernels + C-ABI) and host/ (C++ host
framework mirroring the reference's getFactoriesSingleton()/Filter interface).
"""
from . import _native  # noqa: F401  (loads libb200sdr.so or raises)
from .chain import AM, FM, NONE, Chain, fm_gain  # noqa: F401
from .channelizer import Channelizer  # noqa: F401
from . import ops, sharding, synth, taps  # noqa: F401

__all__ = ["AM", "FM", "NONE", "Chain", "Channelizer", "fm_gain", "ops", "taps", "synth", "sharding"]
