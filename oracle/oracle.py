"""ctypes/numpy front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under ``cuda_sdr_b200/`` imports
it.  See the header of oracle.c for what is pinned by the reference's own tests and what is
"parity unpinned".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

AM, FM, NONE = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        sz, vp, f32, i32, u64 = C.c_size_t, C.c_void_p, C.c_float, C.c_int, C.c_uint64
        L.orc_num_threads.restype = i32
        L.orc_set_num_threads.argtypes = [i32]
        for name in ("orc_fir_num_outputs_verbatim", "orc_fir_num_outputs"):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = sz, [sz, sz, sz]
        L.orc_fm_num_outputs.restype, L.orc_fm_num_outputs.argtypes = sz, [sz]
        L.orc_fm_gain.restype, L.orc_fm_gain.argtypes = f32, [f32, f32]
        L.orc_int8_to_norm_float.argtypes = [vp, vp, sz]
        L.orc_cosine_c.argtypes = [f32, f32, vp, sz]
        L.orc_cosine_f.argtypes = [f32, f32, vp, sz]
        L.orc_cosine_delta.restype, L.orc_cosine_delta.argtypes = f32, [f32, f32]
        L.orc_cosine_phi_end.restype, L.orc_cosine_phi_end.argtypes = f32, [f32, sz, f32]
        L.orc_cosine_next_phi.restype, L.orc_cosine_next_phi.argtypes = f32, [f32]
        L.orc_multiply_cc.argtypes = [vp, vp, vp, sz]
        L.orc_quad_am_demod.argtypes = [vp, vp, sz]
        L.orc_quad_fm_demod.argtypes = [vp, vp, f32, sz]
        L.orc_add_const_ff.argtypes = [vp, f32, vp, sz]
        L.orc_add_to_magnitude.argtypes = [vp, f32, vp, sz]
        for name in ("orc_fir_fc", "orc_fir_ff", "orc_fir_cc", "orc_fir_cf"):
            getattr(L, name).argtypes = [sz, vp, sz, vp, vp, sz]
        L.orc_phase_step.restype, L.orc_phase_step.argtypes = u64, [C.c_double, C.c_double]
        L.orc_chain.restype = sz
        L.orc_chain.argtypes = [vp, i32, sz, u64, i32, u64, vp, sz, sz, i32, f32, vp, sz, sz, vp, vp, vp]
        L.orc_chain_num_outputs.restype, L.orc_chain_num_outputs.argtypes = sz, [sz, sz, sz, i32, sz, sz]
        L.orc_fm_demod_fused.restype = sz
        L.orc_fm_demod_fused.argtypes = [f32, f32, f32, f32, sz, sz, vp, sz, vp, sz, vp]
        _lib = L
    return _lib


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


# ---------------------------------------------------------------------------------------------------
# counts
# ---------------------------------------------------------------------------------------------------
def fir_num_outputs(n_in: int, taps: int, decim: int) -> int:
    return int(lib().orc_fir_num_outputs(n_in, taps, decim))


def fir_num_outputs_verbatim(n_in: int, taps: int, decim: int) -> int:
    return int(lib().orc_fir_num_outputs_verbatim(n_in, taps, decim))


def fm_num_outputs(n_in: int) -> int:
    return int(lib().orc_fm_num_outputs(n_in))


def fm_gain(fs: float, deviation: float) -> float:
    return float(lib().orc_fm_gain(fs, deviation))


def chain_num_outputs(n: int, T1: int, D1: int, modulation: int, T2: int, D2: int) -> int:
    return int(lib().orc_chain_num_outputs(n, T1, D1, modulation, T2, D2))


def phase_step(frequency: float, sample_rate: float) -> int:
    return int(lib().orc_phase_step(float(frequency), float(sample_rate)))


# ---------------------------------------------------------------------------------------------------
# element-wise ops.  Complex arrays are numpy complex (in: complex64, out: complex128).
# ---------------------------------------------------------------------------------------------------
def int8_to_norm_float(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.int8)
    out = np.empty(x.shape, dtype=np.float32)
    lib().orc_int8_to_norm_float(_p(x), _p(out), x.size)
    return out


def cosine_c(phi_start: float, phi_end: float, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.complex128)
    lib().orc_cosine_c(phi_start, phi_end, _p(out), n)
    return out


def cosine_f(phi_start: float, phi_end: float, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    lib().orc_cosine_f(phi_start, phi_end, _p(out), n)
    return out


def multiply_cc(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.complex64)
    b = np.ascontiguousarray(b, dtype=np.complex64)
    out = np.empty(a.size, dtype=np.complex128)
    lib().orc_multiply_cc(_p(a), _p(b), _p(out), a.size)
    return out


def quad_am_demod(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty(x.size, dtype=np.float64)
    lib().orc_quad_am_demod(_p(x), _p(out), x.size)
    return out


def quad_fm_demod(x: np.ndarray, gain: float) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.complex64)
    n = fm_num_outputs(x.size)
    out = np.empty(n, dtype=np.float64)
    lib().orc_quad_fm_demod(_p(x), _p(out), gain, n)
    return out


def add_const_ff(x: np.ndarray, c: float) -> np.ndarray:
    x = _f32(x)
    out = np.empty(x.size, dtype=np.float64)
    lib().orc_add_const_ff(_p(x), c, _p(out), x.size)
    return out


def add_to_magnitude(x: np.ndarray, c: float) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty(x.size, dtype=np.complex128)
    lib().orc_add_to_magnitude(_p(x), c, _p(out), x.size)
    return out


# ---------------------------------------------------------------------------------------------------
# FIR.  kind = "fc" | "ff" | "cc" | "cf"  (tap type, element type) as in Fir.cpp:229-269.
# ---------------------------------------------------------------------------------------------------
_FIR = {
    "fc": ("orc_fir_fc", np.float32, np.complex64, np.complex128),
    "ff": ("orc_fir_ff", np.float32, np.float32, np.float64),
    "cc": ("orc_fir_cc", np.complex64, np.complex64, np.complex128),
    "cf": ("orc_fir_cf", np.complex64, np.float32, np.complex128),
}


def fir(kind: str, taps: np.ndarray, x: np.ndarray, decim: int, n_out: int | None = None) -> np.ndarray:
    fn, tt, xt, ot = _FIR[kind]
    taps = np.ascontiguousarray(taps, dtype=tt)
    x = np.ascontiguousarray(x, dtype=xt)
    decim = max(1, int(decim))
    avail = fir_num_outputs(x.size, taps.size, decim)
    if n_out is None:
        n_out = avail
    assert n_out <= avail or (n_out - 1) * decim + taps.size <= x.size
    out = np.empty(n_out, dtype=ot)
    if n_out:
        getattr(lib(), fn)(decim, _p(taps), taps.size, _p(x), _p(out), n_out)
    return out


# ---------------------------------------------------------------------------------------------------
# the chain
# ---------------------------------------------------------------------------------------------------
@dataclass
class ChainSpec:
    """Parameters of one int8/cf32 -> mix -> FIR -> demod -> audio-FIR channel (reference graph:
    nbfm_test.cpp:256-354, RfToPcmAudioFactory.cpp:214-304)."""

    sample_rate: float
    frequency: float  # cosine-source frequency (tuned - channel), RfToPcmAudioFactory.cpp:225
    taps1: np.ndarray
    decim1: int
    modulation: int = AM
    fm_gain: float = 1.0
    taps2: np.ndarray | None = None
    decim2: int = 1
    mix: bool = True
    input_int8: bool = True
    step: int = field(init=False)

    def __post_init__(self):
        self.taps1 = _f32(self.taps1)
        if self.taps2 is not None:
            self.taps2 = _f32(self.taps2)
        self.step = phase_step(self.frequency, self.sample_rate) if self.mix else 0

    @property
    def T1(self) -> int:
        return int(self.taps1.size)

    @property
    def T2(self) -> int:
        return 0 if self.taps2 is None else int(self.taps2.size)

    def num_outputs(self, n: int) -> int:
        return chain_num_outputs(n, self.T1, self.decim1, self.modulation, self.T2, self.decim2)


def chain(spec: ChainSpec, x: np.ndarray, n0: int = 0, want_rf: bool = False, want_demod: bool = False):
    """Run the fp64 chain over one block.  ``x``: int8 array of interleaved IQ (len 2n) when
    spec.input_int8 else complex64 (len n).  Returns (audio, rf|None, demod|None)."""
    if spec.input_int8:
        x = np.ascontiguousarray(x, dtype=np.int8)
        n = x.size // 2
    else:
        x = np.ascontiguousarray(x, dtype=np.complex64)
        n = x.size
    n_rf = fir_num_outputs(n, spec.T1, spec.decim1)
    n_demod = n_rf if spec.modulation != FM else fm_num_outputs(n_rf)
    n_final = spec.num_outputs(n)
    rf = np.empty(max(n_rf, 1), dtype=np.complex128) if (want_rf or spec.modulation == NONE) else None
    demod = np.empty(max(n_demod, 1), dtype=np.float64) if want_demod and spec.modulation != NONE else None
    audio = np.empty(max(n_final, 1), dtype=np.float64) if spec.modulation != NONE else None
    got = lib().orc_chain(
        _p(x), int(spec.input_int8), n, n0, int(spec.mix), spec.step,
        _p(spec.taps1), spec.T1, spec.decim1, spec.modulation, spec.fm_gain,
        _p(spec.taps2), spec.T2, spec.decim2, _p(rf), _p(demod), _p(audio))
    assert got == n_final, (got, n_final)
    if rf is not None:
        rf = rf[:n_rf]
    if demod is not None:
        demod = demod[:n_demod]
    if audio is not None:
        audio = audio[:n_final]
    if spec.modulation == NONE:
        return rf, rf, None
    return audio, rf, demod


# ---------------------------------------------------------------------------------------------------
# Host-side stream model of one Filter port: accumulates committed input, produces
# min(available, room) outputs per readOutput and consumes exactly nOut*D inputs
# (Fir.cpp:210-278, BaseSink.cpp:61-116,150-170).  Used to check chunking-independence.
# ---------------------------------------------------------------------------------------------------
class FirStreamModel:
    def __init__(self, kind: str, taps: np.ndarray, decim: int):
        self.kind, self.taps, self.decim = kind, np.asarray(taps), max(1, int(decim))
        self.xt = _FIR[kind][2]
        self.pending = np.empty(0, dtype=self.xt)

    def commit(self, x: np.ndarray) -> None:
        self.pending = np.concatenate([self.pending, np.asarray(x, dtype=self.xt)])

    def output_size(self) -> int:
        return fir_num_outputs(self.pending.size, self.taps.size, self.decim)

    def read(self, room: int | None = None) -> np.ndarray:
        n = self.output_size()
        if room is not None:
            n = min(n, room)
        out = fir(self.kind, self.taps, self.pending, self.decim, n)
        self.pending = self.pending[n * self.decim:]
        return out
