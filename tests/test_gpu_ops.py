"""GPU parity of the single-op kernels, called through the gsdr C-ABI, against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.util import REL_TOL, assert_close, assert_fm_close, rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from cuda_sdr_b200 import ops as _ops
    return _ops


def _cplx(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 255, 256, 4099, (1 << 20) + 3])
def test_int8_to_float_bit_exact(ops, n):
    rng = np.random.default_rng(n)
    x = rng.integers(-128, 128, size=n, dtype=np.int8)
    got = ops.int8_to_norm_float(torch.from_numpy(x).to(DEV)).cpu().numpy()
    ref = orc.int8_to_norm_float(x)
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_int8_to_float_all_values_and_unaligned(ops):
    x = np.tile(np.arange(-128, 128, dtype=np.int8), 9)
    d = torch.from_numpy(x).to(DEV)
    for off in (0, 1, 3, 16, 21):
        got = ops.int8_to_norm_float(d[off:].contiguous() if off == 0 else d[off:]).cpu().numpy()
        assert np.array_equal(got, x[off:].astype(np.float32) / np.float32(128))


def test_cosine_kat_through_abi(ops, cosine_kat):
    fs, f = float(cosine_kat["sample_rate"]), float(cosine_kat["frequency"])
    n = 104  # 101 values rounded up to the 32-byte allocator granule (see tests/golden/cosine_kat.json)
    delta = orc.lib().orc_cosine_delta(fs, f)
    phi_end = orc.lib().orc_cosine_phi_end(0.0, n, delta)
    got = ops.cosine_c(0.0, phi_end, n, DEV).cpu().numpy()[: cosine_kat["output_value_count"]]
    i = np.arange(got.size, dtype=np.float32)
    theta = i * np.float32(f) / np.float32(fs) * np.float32(np.pi) * np.float32(2.0)
    assert np.max(np.abs(got.real - np.cos(theta))) < cosine_kat["tolerance"]
    assert np.max(np.abs(got.imag - np.sin(theta))) < cosine_kat["tolerance"]


@pytest.mark.parametrize("n,phi0,delta", [(1, 0.0, 0.1), (101, 0.0, 0.0628), (131072, 1.5, 0.4041), (1000003, 6.2, -2.9)])
def test_cosine_vs_oracle(ops, n, phi0, delta):
    phi_end = float(np.float32(phi0) + np.float32(n) * np.float32(delta))
    assert_close(ops.cosine_c(phi0, phi_end, n, DEV).cpu().numpy(), orc.cosine_c(phi0, phi_end, n), what="cosineC")
    assert_close(ops.cosine_f(phi0, phi_end, n, DEV).cpu().numpy(), orc.cosine_f(phi0, phi_end, n), what="cosineF")


@pytest.mark.parametrize("n", [0, 1, 2, 3, 1001, 1 << 18])
def test_multiply_magnitude_addconst(ops, n):
    rng = np.random.default_rng(100 + n)
    a, b = _cplx(rng, n), _cplx(rng, n)
    da, db = torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV)
    assert_close(ops.multiply_cc(da, db).cpu().numpy(), orc.multiply_cc(a, b), what="multiplyCC")
    assert_close(ops.quad_am_demod(da).cpu().numpy(), orc.quad_am_demod(a), what="quadAm")
    assert_close(ops.magnitude(da).cpu().numpy(), orc.quad_am_demod(a), what="magnitude")
    assert_close(ops.add_to_magnitude(da, 0.25).cpu().numpy(), orc.add_to_magnitude(a, 0.25), what="addToMagnitude")
    r = a.real.copy()
    assert_close(ops.add_const_ff(torch.from_numpy(r).to(DEV), -1.5).cpu().numpy(), orc.add_const_ff(r, -1.5), what="addConst")


@pytest.mark.parametrize("n", [1, 2, 33, 4097, 1 << 18])
def test_quad_fm_demod(ops, n):
    rng = np.random.default_rng(200 + n)
    x = _cplx(rng, n)
    gain = 1.7
    got = ops.quad_fm_demod(torch.from_numpy(x).to(DEV), gain).cpu().numpy()
    assert got.size == orc.fm_num_outputs(n)
    assert_fm_close(got, orc.quad_fm_demod(x, gain), gain, what="quadFm")


# ---- FIR ---------------------------------------------------------------------------------------------
def test_fir_kats_through_abi(ops, fir_kat):
    for case in fir_kat["cases"]:
        taps = torch.tensor(case["taps"], dtype=torch.float32, device=DEV)
        x = np.array([complex(a, b) for c in case["commits"] for a, b in c], dtype=np.complex64)
        pending = torch.from_numpy(x).to(DEV)
        D = case["decimation"]
        for read in case["reads"]:
            n = ops.fir_num_outputs(pending.numel(), taps.numel(), D)
            if read["room_elements"] is not None:
                n = min(n, read["room_elements"])
            got = ops.fir("fc", taps, pending, D, n).cpu().numpy()
            exp = np.array([complex(a, b) for a, b in read["expected"]], dtype=np.complex64)
            assert got.size == exp.size
            assert np.max(np.abs(got - exp)) < fir_kat["tolerance"]
            pending = pending[n * D:]  # Fir.cpp:274-276: consume exactly nOut*D inputs


FIR_SHAPES = [
    # (T, D, n_in)  -- rows path needs D even (cf32) and M = ceil(T/D) <= 8; others exercise the direct path
    (2, 2, 5), (3, 2, 8), (63, 10, 1 << 16), (63, 10, 12345), (101, 40, 1 << 17), (16, 16, 4096), (7, 8, 999),
    (33, 4, 5000), (64, 2, 3000), (129, 1, 2000), (545, 80, 1 << 17), (1025, 3, 9000), (5, 7, 100), (1, 1, 17),
    (2049, 640, 1 << 18),
]


@pytest.mark.parametrize("T,D,n", FIR_SHAPES)
def test_fir_fc_vs_oracle(ops, T, D, n):
    rng = np.random.default_rng(T * 1000 + D)
    taps = (rng.standard_normal(T) / np.sqrt(T)).astype(np.float32)
    x = _cplx(rng, n)
    got = ops.fir("fc", torch.from_numpy(taps).to(DEV), torch.from_numpy(x).to(DEV), D).cpu().numpy()
    ref = orc.fir("fc", taps, x, D)
    assert got.size == ref.size == orc.fir_num_outputs(n, T, D)
    assert_close(got, ref, what=f"firFC T={T} D={D}")


# the corners and the diagonal of bench.py --workload firsweep (T in 32..4096, D in 1..64), every kernel the
# dispatcher picks for them (rows, staged direct, window with 5 / 4 / 2 / 1 phases per pass and one or several tap chunks)
SWEEP_SHAPES = [(4096, 64, 1 << 20), (4096, 1, 30000), (1024, 64, 1 << 20), (32, 64, 1 << 19), (32, 1, 20000), (512, 32, 1 << 19),
                (256, 16, 1 << 18), (128, 8, 1 << 17), (64, 4, 1 << 16), (2048, 2, 40000), (4096, 16, 1 << 19), (1024, 5, 1 << 17),
                # the 16- and 32-partial rows kernels: full (M = 16, 32), padded (M = 19, 15, 9, 25), one and two rows per thread,
                # rows with and without the 16-byte padding (D * 8 a multiple of 64 or not)
                (1024, 32, 1 << 19), (512, 16, 1 << 18), (2048, 64, 1 << 20), (1000, 32, 1 << 19), (600, 32, 1 << 18), (300, 20, 1 << 17),
                (400, 16, 1 << 17), (290, 34, 1 << 17), (1200, 48, 1 << 19)]


@pytest.mark.parametrize("T,D,n", SWEEP_SHAPES)
def test_fir_fc_sweep_cells_vs_oracle(ops, T, D, n):
    rng = np.random.default_rng(T * 131 + D)
    taps = (rng.standard_normal(T) / np.sqrt(T)).astype(np.float32)
    x = _cplx(rng, n + 37)  # ragged length
    got = ops.fir("fc", torch.from_numpy(taps).to(DEV), torch.from_numpy(x).to(DEV), D).cpu().numpy()
    ref = orc.fir("fc", taps, x, D)
    assert got.size == ref.size == orc.fir_num_outputs(x.size, T, D)
    assert_close(got, ref, what=f"firFC sweep cell T={T} D={D}")


@pytest.mark.parametrize("T,D", [(4096, 64), (1024, 64), (4096, 1), (32, 64)])
def test_fir_fc_sweep_cells_at_bench_size(ops, T, D):
    """The sweep's own size (2^26 complex samples): exact output count, and windows of outputs (first, last, across CTA borders,
    random) against the oracle run on exactly the input samples those outputs read (Fir.cpp:262-266)."""
    n = 1 << 26
    g = torch.Generator(device=DEV).manual_seed(T + D)
    xd = torch.view_as_complex(torch.randn(n, 2, device=DEV, dtype=torch.float32, generator=g))
    rng = np.random.default_rng(T + D)
    taps = (rng.standard_normal(T) / np.sqrt(T)).astype(np.float32)
    got = ops.fir("fc", torch.from_numpy(taps).to(DEV), xd, D)
    n_out = orc.fir_num_outputs(n, T, D)
    assert got.numel() == n_out
    W = 96
    starts = [0, n_out - W, 1024 - 40, 1024 * 147 - 50] + [int(v) for v in rng.integers(0, n_out - W, size=6)]
    for k0 in starts:
        xs = xd[k0 * D: (k0 + W) * D + T].cpu().numpy()  # the slice keeps the count rule of the whole input (Fir.cpp:181-186)
        ref = orc.fir("fc", taps, xs, D)[:W]
        assert ref.size == W
        assert_close(got[k0: k0 + W].cpu().numpy(), ref, what=f"firFC T={T} D={D} at 2^26, outputs {k0}..{k0 + W}")


@pytest.mark.parametrize("kind", ["ff", "cc", "cf"])
@pytest.mark.parametrize("T,D,n", [(2, 2, 5), (129, 10, 50000), (273, 5, 40000), (31, 1, 1000), (1500, 4, 20000)])
def test_fir_other_kinds_vs_oracle(ops, kind, T, D, n):
    rng = np.random.default_rng(T + D + len(kind))
    taps = (_cplx(rng, T) / np.sqrt(T)).astype(np.complex64) if kind[0] == "c" else (rng.standard_normal(T) / np.sqrt(T)).astype(np.float32)
    x = _cplx(rng, n) if kind[1] == "c" else rng.standard_normal(n).astype(np.float32)
    got = ops.fir(kind, torch.from_numpy(np.ascontiguousarray(taps)).to(DEV), torch.from_numpy(x).to(DEV), D).cpu().numpy()
    assert_close(got, orc.fir(kind, taps, x, D), what=f"fir{kind.upper()} T={T} D={D}")


def test_fir_fc_unaligned_input_takes_direct_path_and_matches(ops):
    rng = np.random.default_rng(5)
    taps = (rng.standard_normal(63) / 8).astype(np.float32)
    x = _cplx(rng, 70001)
    dx = torch.from_numpy(x).to(DEV)
    for off in (1, 3):  # 8-byte offsets: not 16-byte aligned -> the TMA rows kernel is not eligible
        got = ops.fir("fc", torch.from_numpy(taps).to(DEV), dx[off:], 10).cpu().numpy()
        assert_close(got, orc.fir("fc", taps, x[off:], 10), what=f"unaligned off={off}")


def test_fir_c1_config_counts_and_values(ops):
    """BASELINE configs[0]: 2^20 complex-float samples, 63-tap low-pass, decimate-by-10 -> 104851 outputs."""
    from cuda_sdr_b200 import taps as tapdesign
    rng = np.random.default_rng(0x5D120000)
    x = _cplx(rng, 1 << 20)
    h = tapdesign.lowpass(63, 0.04, 1.0)
    dx, dh = torch.from_numpy(x).to(DEV), torch.from_numpy(h).to(DEV)
    got = ops.fir("fc", dh, dx, 10).cpu().numpy()
    ref = orc.fir("fc", h, x, 10)
    assert got.size == 104851
    assert_close(got, ref, what="C1")
    # chunked (1 MiB steps = 131072 samples, and a prime chunk) must agree bit-for-bit with one shot
    for chunk in (131072, 99991):
        outs, pending = [], dx[:0]
        for s in range(0, x.size, chunk):
            pending = torch.cat([pending, dx[s:s + chunk]])
            n = ops.fir_num_outputs(pending.numel(), 63, 10)
            if pending.data_ptr() % 16:  # keep the fast path eligible like the host framework does
                pending = pending.clone()
            outs.append(ops.fir("fc", dh, pending, 10, n))
            pending = pending[n * 10:].clone()
        chunked = torch.cat(outs).cpu().numpy()
        assert chunked.size == got.size
        assert np.array_equal(chunked.view(np.uint32), got.view(np.uint32)), f"chunk={chunk}"


def test_fm_demod_fused_vs_oracle(ops):
    rng = np.random.default_rng(11)
    from cuda_sdr_b200 import taps as tapdesign
    fs, tuned, chan, dev_hz, D = 2.4e6, 100.0e6, 100.3e6, 75e3, 8
    h = tapdesign.lowpass(61, 100e3, fs)
    n_out = 5000
    n_in = n_out * D + h.size + D  # the reference count rule needs D-1 more inputs than the maths
    t = np.arange(n_in) / fs
    x = (np.exp(2j * np.pi * (300e3 * t + 3.0 * np.sin(2 * np.pi * 1e3 * t))) + 0.05 * _cplx(rng, n_in)).astype(np.complex64)
    got = ops.fm_demod_fused(fs, tuned, chan, dev_hz, D, 777, torch.from_numpy(h).to(DEV), torch.from_numpy(x).to(DEV), n_out).cpu().numpy()
    ref = np.empty(n_out + 8, dtype=np.float64)
    n_ref = orc.lib().orc_fm_demod_fused(fs, tuned, chan, dev_hz, D, 777, h.ctypes.data, h.size, x.ctypes.data, n_in, ref.ctypes.data)
    assert n_ref >= n_out
    gain = (fs / D) / (2 * np.pi * dev_hz)
    assert_fm_close(got, ref[:n_out], gain, what="gsdrFmDemod")
