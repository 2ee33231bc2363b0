"""ctypes binding of libb200sdr.so (the C-ABI declared in include/gsdr/*.h and include/b200sdr/b200sdr.h).

There is deliberately NO fallback: if the shared library is missing or a symbol is absent this module
raises at import time, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libb200sdr.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the DSP kernels.")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

sz, vp, f32, f64, i32, u32, u64 = C.c_size_t, C.c_void_p, C.c_float, C.c_double, C.c_int32, C.c_uint32, C.c_uint64
cudaError = C.c_int
stream_t = C.c_void_p

# name -> (restype, argtypes); one entry per symbol declared in include/gsdr/gsdr.h, conversion.h
GSDR_SYMBOLS = {
    "gsdrInt8ToNormFloat": (cudaError, [vp, vp, sz, i32, stream_t]),
    "gsdrCosineF": (cudaError, [f32, f32, vp, sz, i32, stream_t]),
    "gsdrCosineC": (cudaError, [f32, f32, vp, sz, i32, stream_t]),
    "gsdrMultiplyCC": (cudaError, [vp, vp, vp, sz, i32, stream_t]),
    "gsdrFirFF": (cudaError, [sz, vp, sz, vp, vp, sz, i32, stream_t]),
    "gsdrFirFC": (cudaError, [sz, vp, sz, vp, vp, sz, i32, stream_t]),
    "gsdrFirCC": (cudaError, [sz, vp, sz, vp, vp, sz, i32, stream_t]),
    "gsdrFirCF": (cudaError, [sz, vp, sz, vp, vp, sz, i32, stream_t]),
    "gsdrQuadAmDemod": (cudaError, [vp, vp, sz, i32, stream_t]),
    "gsdrQuadFmDemod": (cudaError, [vp, vp, f32, sz, i32, stream_t]),
    "gsdrMagnitude": (cudaError, [vp, vp, sz, i32, stream_t]),
    "gsdrAddConstFF": (cudaError, [vp, f32, vp, sz, i32, stream_t]),
    "gsdrAddToMagnitude": (cudaError, [vp, f32, vp, sz, i32, stream_t]),
    "gsdrFmDemod": (cudaError, [f32, f32, f32, f32, sz, sz, vp, sz, vp, vp, sz, i32, stream_t]),
}


class ChainConfig(C.Structure):
    """Mirror of b200sdr_chain_config (include/b200sdr/b200sdr.h)."""

    _fields_ = [
        ("struct_size", u32), ("input_type", u32), ("modulation", u32), ("mix", u32),
        ("sample_rate", f64), ("frequency", f64),
        ("rf_taps", C.POINTER(f32)), ("rf_tap_count", sz), ("rf_decimation", sz),
        ("fm_gain", f32), ("reserved0", u32),
        ("audio_taps", C.POINTER(f32)), ("audio_tap_count", sz), ("audio_decimation", sz),
        ("cuda_device", i32), ("reserved1", u32),
    ]


class ChannelizerConfig(C.Structure):
    """Mirror of b200sdr_channelizer_config (include/b200sdr/b200sdr.h)."""

    _fields_ = [
        ("struct_size", u32), ("num_channels", u32), ("sample_rate", f64),
        ("frequencies", C.POINTER(f64)), ("modulations", C.POINTER(u32)), ("fm_gains", C.POINTER(f32)),
        ("rf_taps", C.POINTER(f32)), ("rf_tap_count", sz), ("rf_decimation", sz),
        ("audio_taps", C.POINTER(f32)), ("audio_tap_count", sz), ("audio_decimation", sz),
        ("cuda_device", i32), ("reserved", u32),
    ]


class GatherConfig(C.Structure):
    """Mirror of b200sdr_gather_config (include/b200sdr/b200sdr.h)."""

    _fields_ = [
        ("struct_size", u32), ("rank", i32), ("world", i32), ("slabs", u32), ("cuda_device", i32),
        ("floats_per_rank", C.POINTER(sz)), ("nccl_unique_id", vp), ("mode", u32), ("reserved", u32),
    ]


psz = C.POINTER(sz)
B200SDR_SYMBOLS = {
    "b200sdr_chain_create": (u32, [C.POINTER(ChainConfig), C.POINTER(vp)]),
    "b200sdr_chain_destroy": (None, [vp]),
    "b200sdr_last_error": (C.c_char_p, []),
    "b200sdr_fir_num_outputs": (sz, [sz, sz, sz]),
    "b200sdr_chain_counts": (None, [vp, sz, psz, psz, psz]),
    "b200sdr_chain_input_stride": (sz, [vp]),
    "b200sdr_chain_input_window": (sz, [vp]),
    "b200sdr_phase_step": (u64, [f64, f64]),
    "b200sdr_chain_segment": (u32, [vp, sz, sz, sz, psz, psz, psz, psz]),
    "b200sdr_chain_segment_weighted": (u32, [vp, sz, sz, C.POINTER(f64), sz, psz, psz, psz, psz]),
    "b200sdr_chain_rf_stage": (u32, [vp, vp, sz, u64, vp, sz, stream_t]),
    "b200sdr_chain_audio_stage": (u32, [vp, vp, vp, sz, stream_t]),
    "b200sdr_chain_run": (u32, [vp, vp, sz, u64, vp, vp, sz, stream_t]),
    "b200sdr_chain_process_device": (u32, [vp, vp, sz, u64, vp, vp, sz, psz, stream_t]),
    "b200sdr_chain_process_host": (u32, [vp, vp, sz, u64, vp, sz, psz]),
    "b200sdr_chain_set_host_segment": (u32, [vp, sz]),
    "b200sdr_channelizer_create": (u32, [vp, C.POINTER(vp)]),
    "b200sdr_channelizer_destroy": (None, [vp]),
    "b200sdr_channelizer_counts": (None, [vp, sz, psz, psz]),
    "b200sdr_channelizer_run": (u32, [vp, vp, sz, vp, sz, vp, sz, sz, stream_t]),
    "b200sdr_channelizer_channel_counts": (u32, [vp, u32, sz, psz, psz]),
    "b200sdr_channelizer_process": (u32, [vp, vp, sz, vp, sz, vp, sz, psz, stream_t]),
    "b200sdr_channelizer_variant": (C.c_char_p, [vp]),
    "b200sdr_channelizer_raster": (u32, [C.POINTER(f64), u32, f64, C.POINTER(i32)]),
    "b200sdr_channelizer_segment": (u32, [vp, sz, sz, sz, psz, psz, psz, psz]),
    "b200sdr_nccl_unique_id": (u32, [vp]),
    "b200sdr_gather_create": (u32, [C.POINTER(GatherConfig), C.POINTER(vp)]),
    "b200sdr_gather_destroy": (None, [vp]),
    "b200sdr_gather_exchange_size": (sz, [vp]),
    "b200sdr_gather_export": (u32, [vp, vp]),
    "b200sdr_gather_import": (u32, [vp, vp]),
    "b200sdr_gather_slab": (vp, [vp, u32]),
    "b200sdr_gather_acquire": (u32, [vp, u32, stream_t]),
    "b200sdr_gather_submit": (u32, [vp, u32, psz, stream_t]),
    "b200sdr_gather_finish": (u32, [vp, stream_t]),
    "b200sdr_gather_result": (vp, [vp, u32, i32]),
    "b200sdr_gather_stats": (None, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(i32)]),
    "b200sdr_launch_count": (u64, []),
    "b200sdr_chain_variant": (C.c_char_p, [vp]),
    "b200sdr_version": (C.c_char_p, []),
    "b200sdr_toeplitz_tables": (u32, [C.POINTER(f32), sz, sz, u32, f64, f64, C.POINTER(u32), sz, psz, C.POINTER(f32), C.POINTER(u32), C.POINTER(u32)]),
}

for _name, (_res, _args) in {**GSDR_SYMBOLS, **B200SDR_SYMBOLS}.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library does not export what include/ declares
    _fn.restype = _res
    _fn.argtypes = _args

STATUS_NAMES = ["Success", "UnknownError", "OutOfMemory", "RuntimeError", "InvalidArgument", "InvalidState",
                "OutOfRange", "TimedOut", "NotFound", "ParseError"]  # reference Status.h:22-34


class NativeError(RuntimeError):
    pass


def check_status(status: int, what: str) -> None:
    if status != 0:
        name = STATUS_NAMES[status] if status < len(STATUS_NAMES) else str(status)
        raise NativeError(f"{what}: Status_{name}: {lib.b200sdr_last_error().decode()}")


def check_cuda(err: int, what: str) -> None:
    if err != 0:
        raise NativeError(f"{what}: cudaError {err}")


def launch_count() -> int:
    return int(lib.b200sdr_launch_count())
