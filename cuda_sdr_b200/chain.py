"""Python front-end of the fused chain (include/b200sdr/b200sdr.h).  One `Chain` is one
int8/cf32 -> mix -> decimating FIR -> AM/FM demod -> audio FIR channel, i.e. the node the reference
assembles in RfToPcmAudioFactory.cpp:214-304 (+ Int8ToFloat)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _native as N

AM, FM, NONE = 0, 1, 2
_lib = N.lib


def fm_gain(input_sample_rate: float, fsk_deviation: float) -> float:
    """QuadDemodFactory.h:108-110, evaluated in float32 like the reference."""
    f = np.float32
    return float(f(input_sample_rate) / (f(2.0) * f(math.pi) * f(fsk_deviation) * f(5)))


class Chain:
    def __init__(self, sample_rate: float, frequency: float, rf_taps, rf_decim: int, modulation: int = AM,
                 fm_gain: float = 1.0, audio_taps=None, audio_decim: int = 1, input_int8: bool = True,
                 mix: bool = True, device: int = 0):
        self._rf_taps = np.ascontiguousarray(rf_taps, dtype=np.float32)
        self._audio_taps = None if audio_taps is None else np.ascontiguousarray(audio_taps, dtype=np.float32)
        cfg = N.ChainConfig()
        cfg.struct_size = C.sizeof(N.ChainConfig)
        cfg.input_type = 2 if input_int8 else 0
        cfg.modulation = modulation
        cfg.mix = 1 if mix else 0
        cfg.sample_rate = float(sample_rate)
        cfg.frequency = float(frequency)
        cfg.rf_taps = self._rf_taps.ctypes.data_as(C.POINTER(C.c_float))
        cfg.rf_tap_count = self._rf_taps.size
        cfg.rf_decimation = rf_decim
        cfg.fm_gain = fm_gain
        if self._audio_taps is not None:
            cfg.audio_taps = self._audio_taps.ctypes.data_as(C.POINTER(C.c_float))
            cfg.audio_tap_count = self._audio_taps.size
        cfg.audio_decimation = audio_decim
        cfg.cuda_device = device
        handle = C.c_void_p()
        N.check_status(_lib.b200sdr_chain_create(C.byref(cfg), C.byref(handle)), "b200sdr_chain_create")
        self._h = handle
        self.device = torch.device("cuda", device)
        self.input_int8 = input_int8
        self.modulation = modulation
        self.rf_decim, self.audio_decim = max(1, rf_decim), max(1, audio_decim)
        self.T1 = self._rf_taps.size
        self.T2 = 0 if self._audio_taps is None else self._audio_taps.size

    def close(self):
        if getattr(self, "_h", None):
            _lib.b200sdr_chain_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host arithmetic ---------------------------------------------------------------------------
    @property
    def variant(self) -> str:
        return _lib.b200sdr_chain_variant(self._h).decode()

    @property
    def stride(self) -> int:
        return int(_lib.b200sdr_chain_input_stride(self._h))

    @property
    def window(self) -> int:
        return int(_lib.b200sdr_chain_input_window(self._h))

    def counts(self, n_in: int):
        rf, demod, audio = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _lib.b200sdr_chain_counts(self._h, n_in, C.byref(rf), C.byref(demod), C.byref(audio))
        return rf.value, demod.value, audio.value

    def segment(self, num_audio: int, parts: int, index: int):
        """(first_output, output_count, first_input, input_count) of time-segment `index` of `parts`."""
        v = [C.c_size_t() for _ in range(4)]
        N.check_status(_lib.b200sdr_chain_segment(self._h, num_audio, parts, index, *[C.byref(x) for x in v]),
                       "b200sdr_chain_segment")
        return tuple(x.value for x in v)

    def segment_weighted(self, num_audio: int, weights, index: int):
        """As segment(), with the outputs split in proportion to `weights` (one positive number per part)."""
        w = (C.c_double * len(weights))(*[float(x) for x in weights])
        v = [C.c_size_t() for _ in range(4)]
        N.check_status(_lib.b200sdr_chain_segment_weighted(self._h, num_audio, len(weights), w, index, *[C.byref(x) for x in v]),
                       "b200sdr_chain_segment_weighted")
        return tuple(x.value for x in v)

    def set_host_segment(self, input_samples: int):
        N.check_status(_lib.b200sdr_chain_set_host_segment(self._h, input_samples), "set_host_segment")

    # ---- device path -------------------------------------------------------------------------------
    def _num_inputs(self, x: torch.Tensor) -> int:
        if self.input_int8:
            assert x.dtype == torch.int8
            return x.numel() // 2
        assert x.dtype == torch.complex64
        return x.numel()

    def _out_dtype(self):
        return torch.complex64 if self.modulation == NONE else torch.float32

    def rf_stage(self, x: torch.Tensor, n_out: int, first_index: int = 0, out: torch.Tensor | None = None,
                 n_in: int | None = None) -> torch.Tensor:
        assert x.is_cuda and x.is_contiguous()
        n_in = self._num_inputs(x) if n_in is None else n_in
        out = torch.empty(n_out, dtype=self._out_dtype(), device=x.device) if out is None else out
        st = _lib.b200sdr_chain_rf_stage(self._h, x.data_ptr(), n_in, first_index, out.data_ptr(), n_out,
                                         torch.cuda.current_stream(x.device).cuda_stream)
        N.check_status(st, "b200sdr_chain_rf_stage")
        return out

    def audio_stage(self, demod: torch.Tensor, n_audio: int, out: torch.Tensor | None = None) -> torch.Tensor:
        out = torch.empty(n_audio, dtype=torch.float32, device=demod.device) if out is None else out
        st = _lib.b200sdr_chain_audio_stage(self._h, demod.data_ptr(), out.data_ptr(), n_audio,
                                            torch.cuda.current_stream(demod.device).cuda_stream)
        N.check_status(st, "b200sdr_chain_audio_stage")
        return out

    def run(self, x: torch.Tensor, n_audio: int, first_index: int = 0, out: torch.Tensor | None = None,
            scratch: torch.Tensor | None = None, n_in: int | None = None) -> torch.Tensor:
        """The whole chain producing exactly n_audio outputs (one persistent kernel when the shape allows)."""
        assert x.is_cuda and x.is_contiguous()
        n_in = self._num_inputs(x) if n_in is None else n_in
        out = torch.empty(max(n_audio, 1), dtype=self._out_dtype(), device=x.device) if out is None else out
        if scratch is None and self.T2 and self.modulation != NONE and not self.fused:
            scratch = torch.empty(max((n_audio - 1) * self.audio_decim + self.T2, 1), dtype=torch.float32, device=x.device)
        st = _lib.b200sdr_chain_run(self._h, x.data_ptr(), n_in, first_index, None if scratch is None else scratch.data_ptr(),
                                    out.data_ptr(), n_audio, torch.cuda.current_stream(x.device).cuda_stream)
        N.check_status(st, "b200sdr_chain_run")
        return out[:n_audio]

    @property
    def fused(self) -> bool:
        return self.variant.startswith(("chain<", "toeplitz<"))

    def process_device(self, x: torch.Tensor, first_index: int = 0, out: torch.Tensor | None = None,
                       scratch: torch.Tensor | None = None) -> torch.Tensor:
        """K1 + K2 over one device-resident block; returns the final outputs (a view of `out`)."""
        assert x.is_cuda and x.is_contiguous()
        n_in = self._num_inputs(x)
        _, n_demod, n_audio = self.counts(n_in)
        if out is None:
            out = torch.empty(max(n_audio, 1), dtype=self._out_dtype(), device=x.device)
        if scratch is None and self.T2 and self.modulation != NONE:
            scratch = torch.empty(max(n_demod, 1), dtype=torch.float32, device=x.device)
        got = C.c_size_t()
        st = _lib.b200sdr_chain_process_device(
            self._h, x.data_ptr(), n_in, first_index, None if scratch is None else scratch.data_ptr(), out.data_ptr(),
            out.numel(), C.byref(got), torch.cuda.current_stream(x.device).cuda_stream)
        N.check_status(st, "b200sdr_chain_process_device")
        return out[: got.value]

    # ---- host path ---------------------------------------------------------------------------------
    def process_host(self, x: torch.Tensor, first_index: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
        """Host (ideally pinned) input -> host output through the C-ABI; synchronous."""
        assert not x.is_cuda and x.is_contiguous()
        n_in = self._num_inputs(x)
        _, _, n_audio = self.counts(n_in)
        if out is None:
            out = torch.empty(max(n_audio, 1), dtype=self._out_dtype()).pin_memory()
        got = C.c_size_t()
        st = _lib.b200sdr_chain_process_host(self._h, x.data_ptr(), n_in, first_index, out.data_ptr(), out.numel(), C.byref(got))
        N.check_status(st, "b200sdr_chain_process_host")
        return out[: got.value]
