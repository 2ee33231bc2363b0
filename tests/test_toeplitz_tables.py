"""CPU check of the host-side tables of the Toeplitz chain kernel (csrc/toeplitz.cu: buildToeplitzFragments) through the
C-ABI hook b200sdr_toeplitz_tables: rebuild B from the fragment words exactly as mma.sync.m16n8k32 reads them, run the
int8 contraction in numpy (exact integers), and compare |y| / the FM discriminator with the fp64 oracle.  No GPU: this
pins the algebra (Toeplitz view, mixer folded into the taps, three-digit fixed point, fragment order) that toepKernel
executes; tests/test_gpu_chain.py pins the kernel itself."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc


@pytest.fixture(scope="module")
def sdr():
    import cuda_sdr_b200 as m
    return m


def tables(sdr, taps, D1, freq, fs, mix=True):
    lib = sdr._native.lib
    taps = np.ascontiguousarray(taps, dtype=np.float32)
    words, ks, magic = C.c_size_t(), C.c_uint32(), C.c_uint32()
    tp = taps.ctypes.data_as(C.POINTER(C.c_float))
    st = lib.b200sdr_toeplitz_tables(tp, taps.size, D1, int(mix), freq, fs, None, 0, C.byref(words), None, C.byref(ks), None)
    assert st == 0
    frag = np.zeros(words.value, dtype=np.uint32)
    scale = (C.c_float * 3)()
    st = lib.b200sdr_toeplitz_tables(tp, taps.size, D1, int(mix), freq, fs, frag.ctypes.data_as(C.POINTER(C.c_uint32)), frag.size,
                                     C.byref(words), scale, C.byref(ks), C.byref(magic))
    assert st == 0
    return frag, np.array(list(scale), dtype=np.float64), ks.value, bool(magic.value)


def digits_from_fragments(frag, ks):
    """B_d[k][col] (d = 0..2, k < 64*Q, col < 8) from the words in kernel order: pair q, lane = 4g + t,
    word (ksub*3 + d)*2 + half holds k = 32*(2q + ksub) + 16*half + 4t + e (e = byte 0..3) of column g."""
    Q = (ks + 1) // 2
    w = frag.reshape(Q, 32, 2, 3, 2)  # q, lane, ksub, d, half
    B = np.zeros((3, 64 * Q, 8), dtype=np.int64)
    for q in range(Q):
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for ksub in range(2):
                for d in range(3):
                    for half in range(2):
                        word = int(w[q, lane, ksub, d, half])
                        for e in range(4):
                            b = (word >> (8 * e)) & 0xFF
                            B[d, 32 * (2 * q + ksub) + 16 * half + 4 * t + e, g] = b - 256 if b >= 128 else b
    return B


def contraction(x_bytes, B, scale, D1, n_out):
    """y[4i + n] = sum_k A[i][k] (B[k][2n] + j B[k][2n+1]), A[i][k] = byte 8*D1*i + k of the stream; exact integers per digit."""
    K = B.shape[1]
    rows = (n_out + 3) // 4
    A = np.lib.stride_tricks.as_strided(x_bytes, shape=(rows, K), strides=(8 * D1, 1)).astype(np.int64)
    acc = np.einsum("ik,dkc->dic", A, B)                      # int64, exact
    v = (acc * scale[:, None, None]).sum(axis=0)              # rows x 8
    y = (v[:, 0::2] + 1j * v[:, 1::2]).reshape(-1)            # output 4i + n
    return y[:n_out], acc


@pytest.mark.parametrize("T1,D1,freq,mod", [(101, 40, -1.234e6, orc.AM), (101, 40, -1.234e6, orc.FM), (545, 80, 2.5e6, orc.FM),
                                            (17, 64, 0.3e6, orc.AM), (3, 8, -0.7e6, orc.AM), (40, 40, 0.0, orc.AM)])
def test_toeplitz_contraction_matches_the_oracle(sdr, T1, D1, freq, mod):
    fs = 19.2e6
    taps = sdr.taps.lowpass(T1, 0.45 * fs / D1, fs)
    frag, scale, ks, magic = tables(sdr, taps, D1, freq, fs)
    assert ks == (2 * T1 + 6 * D1 + 31) // 32
    B = digits_from_fragments(frag, ks)
    # rows beyond the taps' support are zero: the kernel may read (and multiply) bytes there
    assert not B[:, 2 * T1 + 6 * D1:, :].any()

    n = 40 * D1 + T1 + 4 * D1 + 7
    x = sdr.synth.int8_iq(n, seed=T1 + D1)
    gain = 0.37
    t2 = np.array([1.0], dtype=np.float32)  # identity audio stage: the demodulated stream is the output
    spec = orc.ChainSpec(fs, freq, taps, D1, mod, gain, t2, 1)
    _, rf, demod = orc.chain(spec, x, n0=0, want_rf=True, want_demod=True)
    n_rf = rf.size
    xb = np.concatenate([x.view(np.uint8).astype(np.int16).astype(np.int8), np.zeros(64 * ((ks + 1) // 2) + 8 * D1 * 4, dtype=np.int8)])
    y, acc = contraction(xb, B, scale, D1, n_rf)

    # the per-output carrier exp(j*w*k*D) is dropped by the kernel: |y| and the discriminator do not see it
    w = 2.0 * np.pi * freq / fs
    carrier = np.exp(1j * w * D1 * np.arange(n_rf))
    err_rf = np.max(np.abs(y * carrier - rf)) / np.max(np.abs(rf))
    assert err_rf < 2e-6, err_rf  # 24-bit fixed-point taps; the oracle's phase is 64-bit fixed-point turns of the same frequency
    if mod == orc.AM:
        got = np.abs(y)
        assert np.max(np.abs(got - demod)) / np.max(np.abs(demod)) < 1e-6
    else:
        d = y[1:] * np.conj(y[:-1]) * np.exp(1j * w * D1)
        got = gain * np.angle(d)
        diff = np.abs(got - demod[: got.size])
        diff = np.minimum(diff, np.abs(2 * np.pi * gain - diff))
        assert np.max(diff) / (np.pi * gain) < 1e-5

    # magic-number accumulators are only legal while every digit sum stays below 2^22 for ANY int8 input
    worst = np.abs(B).sum(axis=1).max() * 128
    assert magic == (worst < 2 ** 22)
    assert np.abs(acc).max() < 2 ** 31


def test_toeplitz_digits_reconstruct_the_complex_taps(sdr):
    fs, D1, T1, freq = 19.2e6, 40, 101, -1.234e6
    taps = sdr.taps.lowpass(T1, 0.45 * fs / D1, fs)
    frag, scale, ks, _ = tables(sdr, taps, D1, freq, fs)
    B = digits_from_fragments(frag, ks)
    val = (B * scale[:, None, None]).sum(axis=0)
    step = sdr._native.lib.b200sdr_phase_step(freq, fs)
    j = np.arange(T1, dtype=object)
    turns = np.array([((int(step) * int(k)) % (1 << 64)) / float(1 << 64) for k in j])
    c = taps.astype(np.float64) * np.exp(2j * np.pi * turns) / 128.0
    tol = 2.0 ** -23 * np.max(np.abs(c))
    for n in range(4):
        rows = 2 * (D1 * n + np.arange(T1))
        assert np.max(np.abs(val[rows, 2 * n] - c.real)) <= tol
        assert np.max(np.abs(val[rows + 1, 2 * n] + c.imag)) <= tol
        assert np.max(np.abs(val[rows, 2 * n + 1] - c.imag)) <= tol
        assert np.max(np.abs(val[rows + 1, 2 * n + 1] - c.real)) <= tol


def test_toeplitz_tables_reject_unsupported_decimation(sdr):
    taps = np.ones(5, dtype=np.float32)
    words = C.c_size_t()
    st = sdr._native.lib.b200sdr_toeplitz_tables(taps.ctypes.data_as(C.POINTER(C.c_float)), 5, 7, 1, 1.0, 10.0, None, 0, C.byref(words), None,
                                                 None, None)
    assert st != 0
