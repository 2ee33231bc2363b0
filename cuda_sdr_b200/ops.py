"""Thin torch-tensor front-ends of the gsdr C-ABI (include/gsdr/gsdr.h).  Each function enqueues exactly
one call on the current CUDA stream of the tensor's device; pointers are passed through untouched."""
from __future__ import annotations

import torch

from . import _native as N

_lib = N.lib


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _dev(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise ValueError("the DSP kernels have no CPU path: tensors must live on a CUDA device")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def int8_to_norm_float(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    assert x.dtype == torch.int8 and x.is_contiguous()
    out = torch.empty(x.numel(), dtype=torch.float32, device=x.device) if out is None else out
    N.check_cuda(_lib.gsdrInt8ToNormFloat(x.data_ptr(), out.data_ptr(), x.numel(), _dev(x), _stream(x)), "gsdrInt8ToNormFloat")
    return out


def cosine_c(phi_start: float, phi_end: float, n: int, device) -> torch.Tensor:
    out = torch.empty(n, dtype=torch.complex64, device=device)
    N.check_cuda(_lib.gsdrCosineC(phi_start, phi_end, out.data_ptr(), n, _dev(out), _stream(out)), "gsdrCosineC")
    return out


def cosine_f(phi_start: float, phi_end: float, n: int, device) -> torch.Tensor:
    out = torch.empty(n, dtype=torch.float32, device=device)
    N.check_cuda(_lib.gsdrCosineF(phi_start, phi_end, out.data_ptr(), n, _dev(out), _stream(out)), "gsdrCosineF")
    return out


def multiply_cc(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    assert a.dtype == b.dtype == torch.complex64 and a.numel() == b.numel()
    out = torch.empty_like(a)
    N.check_cuda(_lib.gsdrMultiplyCC(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _dev(a), _stream(a)), "gsdrMultiplyCC")
    return out


_FIR = {
    "ff": ("gsdrFirFF", torch.float32, torch.float32, torch.float32),
    "fc": ("gsdrFirFC", torch.float32, torch.complex64, torch.complex64),
    "cc": ("gsdrFirCC", torch.complex64, torch.complex64, torch.complex64),
    "cf": ("gsdrFirCF", torch.complex64, torch.float32, torch.complex64),
}


def fir_num_outputs(n_in: int, taps: int, decim: int) -> int:
    return int(_lib.b200sdr_fir_num_outputs(n_in, taps, decim))


def fir(kind: str, taps: torch.Tensor, x: torch.Tensor, decim: int, n_out: int | None = None) -> torch.Tensor:
    fn, tt, xt, ot = _FIR[kind]
    assert taps.dtype == tt and x.dtype == xt and taps.is_cuda and x.is_cuda
    if n_out is None:
        n_out = fir_num_outputs(x.numel(), taps.numel(), decim)
    assert n_out == 0 or (n_out - 1) * max(1, decim) + taps.numel() <= x.numel()
    out = torch.empty(n_out, dtype=ot, device=x.device)
    N.check_cuda(getattr(_lib, fn)(decim, taps.data_ptr(), taps.numel(), x.data_ptr(), out.data_ptr(), n_out, _dev(x), _stream(x)), fn)
    return out


def quad_am_demod(x: torch.Tensor) -> torch.Tensor:
    assert x.dtype == torch.complex64
    out = torch.empty(x.numel(), dtype=torch.float32, device=x.device)
    N.check_cuda(_lib.gsdrQuadAmDemod(x.data_ptr(), out.data_ptr(), x.numel(), _dev(x), _stream(x)), "gsdrQuadAmDemod")
    return out


def magnitude(x: torch.Tensor) -> torch.Tensor:
    assert x.dtype == torch.complex64
    out = torch.empty(x.numel(), dtype=torch.float32, device=x.device)
    N.check_cuda(_lib.gsdrMagnitude(x.data_ptr(), out.data_ptr(), x.numel(), _dev(x), _stream(x)), "gsdrMagnitude")
    return out


def quad_fm_demod(x: torch.Tensor, gain: float) -> torch.Tensor:
    assert x.dtype == torch.complex64
    n = max(0, x.numel() - 1)
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    N.check_cuda(_lib.gsdrQuadFmDemod(x.data_ptr(), out.data_ptr(), gain, n, _dev(x), _stream(x)), "gsdrQuadFmDemod")
    return out


def add_const_ff(x: torch.Tensor, c: float) -> torch.Tensor:
    assert x.dtype == torch.float32
    out = torch.empty_like(x)
    N.check_cuda(_lib.gsdrAddConstFF(x.data_ptr(), c, out.data_ptr(), x.numel(), _dev(x), _stream(x)), "gsdrAddConstFF")
    return out


def add_to_magnitude(x: torch.Tensor, c: float) -> torch.Tensor:
    assert x.dtype == torch.complex64
    out = torch.empty_like(x)
    N.check_cuda(_lib.gsdrAddToMagnitude(x.data_ptr(), c, out.data_ptr(), x.numel(), _dev(x), _stream(x)), "gsdrAddToMagnitude")
    return out


def fm_demod_fused(rf_rate: float, tuned: float, channel: float, deviation: float, decim: int, first_offset: int,
                   taps: torch.Tensor, x: torch.Tensor, n_out: int) -> torch.Tensor:
    assert taps.dtype == torch.float32 and x.dtype == torch.complex64
    assert n_out * decim + taps.numel() <= x.numel()
    out = torch.empty(n_out, dtype=torch.float32, device=x.device)
    N.check_cuda(_lib.gsdrFmDemod(rf_rate, tuned, channel, deviation, decim, first_offset, taps.data_ptr(), taps.numel(),
                                  x.data_ptr(), out.data_ptr(), n_out, _dev(x), _stream(x)), "gsdrFmDemod")
    return out
