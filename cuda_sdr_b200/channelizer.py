"""Python front-end of the wideband channelizer (include/b200sdr/b200sdr.h, BASELINE config C5): many AM/FM channels
out of one int8 IQ stream.  Sharding over GPUs is by channel: every rank builds a Channelizer over its own channels
(`sharding.channels_of_rank`), the input is replicated, nothing is exchanged on the filter path."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N

_lib = N.lib


class Channelizer:
    def __init__(self, sample_rate: float, frequencies, modulations, rf_taps, rf_decim: int, audio_taps, audio_decim: int,
                 fm_gains=None, device: int = 0):
        self._freq = np.ascontiguousarray(frequencies, dtype=np.float64)
        self._mod = np.ascontiguousarray(modulations, dtype=np.uint32)
        self._gain = np.ascontiguousarray(np.ones(self._freq.size) if fm_gains is None else fm_gains, dtype=np.float32)
        self._t1 = np.ascontiguousarray(rf_taps, dtype=np.float32)
        self._t2 = np.ascontiguousarray(audio_taps, dtype=np.float32)
        assert self._freq.size == self._mod.size == self._gain.size
        cfg = N.ChannelizerConfig()
        cfg.struct_size = C.sizeof(N.ChannelizerConfig)
        cfg.num_channels = self._freq.size
        cfg.sample_rate = float(sample_rate)
        cfg.frequencies = self._freq.ctypes.data_as(C.POINTER(C.c_double))
        cfg.modulations = self._mod.ctypes.data_as(C.POINTER(C.c_uint32))
        cfg.fm_gains = self._gain.ctypes.data_as(C.POINTER(C.c_float))
        cfg.rf_taps = self._t1.ctypes.data_as(C.POINTER(C.c_float))
        cfg.rf_tap_count = self._t1.size
        cfg.rf_decimation = rf_decim
        cfg.audio_taps = self._t2.ctypes.data_as(C.POINTER(C.c_float))
        cfg.audio_tap_count = self._t2.size
        cfg.audio_decimation = audio_decim
        cfg.cuda_device = device
        handle = C.c_void_p()
        N.check_status(_lib.b200sdr_channelizer_create(C.byref(cfg), C.byref(handle)), "b200sdr_channelizer_create")
        self._h = handle
        self.num_channels = int(self._freq.size)
        self.T2, self.D2 = int(self._t2.size), max(1, int(audio_decim))
        self.device = torch.device("cuda", device)

    def close(self):
        if getattr(self, "_h", None):
            _lib.b200sdr_channelizer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def variant(self) -> str:
        return _lib.b200sdr_channelizer_variant(self._h).decode()

    def counts(self, n_in: int):
        demod, audio = C.c_size_t(), C.c_size_t()
        _lib.b200sdr_channelizer_counts(self._h, n_in, C.byref(demod), C.byref(audio))
        return demod.value, audio.value

    def segment(self, num_audio: int, parts: int, index: int):
        """(first_output, output_count, first_input, input_count) of time-segment `index` of `parts` (all channels per segment)."""
        v = [C.c_size_t() for _ in range(4)]
        N.check_status(_lib.b200sdr_channelizer_segment(self._h, num_audio, parts, index, *[C.byref(x) for x in v]),
                       "b200sdr_channelizer_segment")
        return tuple(x.value for x in v)

    def channel_counts(self, channel: int, n_in: int):
        """(demod, audio) counts of one channel by the reference's per-node rules (an AM channel keeps its own count)."""
        demod, audio = C.c_size_t(), C.c_size_t()
        N.check_status(_lib.b200sdr_channelizer_channel_counts(self._h, channel, n_in, C.byref(demod), C.byref(audio)),
                       "b200sdr_channelizer_channel_counts")
        return demod.value, audio.value

    def process(self, x: torch.Tensor, out: torch.Tensor | None = None, scratch: torch.Tensor | None = None):
        """One block by the reference's count rules.  Returns (audio[num_channels, max count], counts[num_channels])."""
        assert x.is_cuda and x.dtype == torch.int8 and x.is_contiguous()
        n_in = x.numel() // 2
        n_max = max(self.channel_counts(c, n_in)[1] for c in range(self.num_channels))
        n_demod = max((n_max - 1) * self.D2 + self.T2 if n_max else 0, self.T2)
        if out is None:
            out = torch.zeros(self.num_channels, max(n_max, 1), dtype=torch.float32, device=x.device)
        if scratch is None:
            # rows 16-byte aligned: the audio stage then copies its tiles with 16-byte loads for every channel
            scratch = torch.empty(self.num_channels, (max(n_demod, 1) + 3) // 4 * 4, dtype=torch.float32, device=x.device)[:, :max(n_demod, 1)]
        counts = (C.c_size_t * self.num_channels)()
        st = _lib.b200sdr_channelizer_process(self._h, x.data_ptr(), n_in, scratch.data_ptr(), scratch.stride(0), out.data_ptr(), out.stride(0),
                                              counts, torch.cuda.current_stream(x.device).cuda_stream)
        N.check_status(st, "b200sdr_channelizer_process")
        return out[:, :n_max], [int(v) for v in counts]

    def run(self, x: torch.Tensor, n_audio: int | None = None, out: torch.Tensor | None = None,
            scratch: torch.Tensor | None = None) -> torch.Tensor:
        """x: device int8 IQ (2 * n_in bytes).  Returns audio[num_channels, n_audio] (device)."""
        assert x.is_cuda and x.dtype == torch.int8 and x.is_contiguous()
        n_in = x.numel() // 2
        if n_audio is None:
            n_audio = self.counts(n_in)[1]
        n_demod = (n_audio - 1) * self.D2 + self.T2 if n_audio else 0
        if out is None:
            out = torch.empty(self.num_channels, max(n_audio, 1), dtype=torch.float32, device=x.device)
        if scratch is None:
            # rows 16-byte aligned: the audio stage then copies its tiles with 16-byte loads for every channel
            scratch = torch.empty(self.num_channels, (max(n_demod, 1) + 3) // 4 * 4, dtype=torch.float32, device=x.device)[:, :max(n_demod, 1)]
        assert out.stride(1) == 1 and scratch.stride(1) == 1
        st = _lib.b200sdr_channelizer_run(self._h, x.data_ptr(), n_in, scratch.data_ptr(), scratch.stride(0), out.data_ptr(), out.stride(0),
                                          n_audio, torch.cuda.current_stream(x.device).cuda_stream)
        N.check_status(st, "b200sdr_channelizer_run")
        return out[:, :n_audio]
