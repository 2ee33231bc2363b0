"""CPU check of the algebra behind pfbKernel (csrc/pfb_kernels.cuh): for channels on a raster fs/N with a common offset,
polyphase filter bank + inverse FFT (the in-place Stockham radix-4/2 schedule the kernel runs, restated in numpy) equals the
per-channel mix -> FIR -> decimate chain of the fp64 oracle, including the 64-bit fixed-point phase steps and the magic-
number sample encoding X = 2^20 + (x + 128) with its precomputed correction.  The kernel itself is pinned on the GPU by
tests/test_gpu_channelizer.py."""
import numpy as np
import pytest

from oracle import oracle as orc


def stockham_inverse(u):
    """Y[c] = sum_r u[r] exp(+2 pi i c r / N): autosort passes of radix 4, then one of radix 2 if log2 N is odd."""
    n = u.size
    a = u.astype(np.complex128).copy()
    ns = 1
    while ns < n:
        radix = 4 if ns * 4 <= n else 2
        b = np.empty_like(a)
        for j in range(n // radix):
            kk = j % ns
            w1 = np.exp(2j * np.pi * kk / (ns * radix))
            v = [a[j + r * (n // radix)] * w1 ** r for r in range(radix)]
            if radix == 4:
                t0, t1, t2, t3 = v[0] + v[2], v[0] - v[2], v[1] + v[3], (v[1] - v[3]) * 1j
                o = [t0 + t2, t1 + t3, t0 - t2, t1 - t3]
            else:
                o = [v[0] + v[1], v[0] - v[1]]
            j0 = (j // ns) * ns * radix + kk
            for r in range(radix):
                b[j0 + r * ns] = o[r]
        a, ns = b, ns * radix
    return a


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128, 256])
def test_stockham_schedule_is_the_inverse_dft(n):
    rng = np.random.default_rng(n)
    u = rng.normal(size=n) + 1j * rng.normal(size=n)
    assert np.max(np.abs(stockham_inverse(u) - np.fft.ifft(u) * n)) < 1e-12 * n


def test_filter_bank_equals_the_per_channel_chain():
    import cuda_sdr_b200 as sdr

    fs, n_fft, T1, D1, f0 = 1.024e6, 16, 400, 64, 7e3
    bins = [0, 3, 15, 9]
    taps = sdr.taps.lowpass(T1, 0.4 * fs / D1, fs)
    n = 6000
    x = sdr.synth.int8_iq(n, seed=5, sample_rate=fs)
    step0 = orc.phase_step(f0, fs)
    turns = np.array([((step0 * j) % (1 << 64)) / float(1 << 64) for j in range(T1)])
    hp = np.zeros(-(-T1 // n_fft) * n_fft, dtype=np.complex128)
    hp[:T1] = taps.astype(np.float64) * np.exp(2j * np.pi * turns) / 128.0
    qn = hp.size // n_fft
    # samples as the kernel feeds them to the multiply-adds, and the correction that removes the offset
    magic = 2.0 ** 20 + 128.0
    xb = x.astype(np.float64).reshape(-1, 2) + magic            # X = 2^20 + (x + 128), exact in fp64
    acc0 = -magic * (1 + 1j) * hp.reshape(qn, n_fft).sum(axis=0)
    n_rf = (n + 1 - T1) // D1
    Y = np.empty((n_rf, n_fft), dtype=np.complex128)
    for k in range(n_rf):
        seg = xb[k * D1: k * D1 + hp.size]
        seg = np.vstack([seg, np.full((hp.size - seg.shape[0], 2), magic)]) if seg.shape[0] < hp.size else seg
        z = seg[:, 0] + 1j * seg[:, 1]
        u = acc0 + (hp * z).reshape(qn, n_fft).sum(axis=0)
        Y[k] = stockham_inverse(u)
    for b in bins:
        f = f0 + b * fs / n_fft
        spec = orc.ChainSpec(fs, f, taps, D1, orc.AM, 1.0, np.array([1.0], dtype=np.float32), 1)
        _, rf, demod = orc.chain(spec, x, want_rf=True, want_demod=True)
        stepc = orc.phase_step(f, fs)
        carrier = np.exp(2j * np.pi * np.array([((stepc * (k * D1)) % (1 << 64)) / float(1 << 64) for k in range(n_rf)]))
        assert np.max(np.abs(Y[:, b] * carrier - rf[:n_rf])) / np.max(np.abs(rf)) < 1e-9
        assert np.max(np.abs(np.abs(Y[:, b]) - demod[:n_rf])) / np.max(demod) < 1e-9
