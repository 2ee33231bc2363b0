"""Pin the CPU oracle against every known-answer vector the reference's own tests hold for the path
(SURVEY.md section 8(c)): tests/FirTests.cpp (two cases) and tests/CosineSourceTests.cpp."""
import numpy as np
import pytest

from oracle import oracle as orc


def _c(lst):
    return np.array([complex(a, b) for a, b in lst], dtype=np.complex64)


def test_fir_kats_through_stream_model(fir_kat):
    for case in fir_kat["cases"]:
        model = orc.FirStreamModel("fc", np.array(case["taps"], np.float32), case["decimation"])
        for commit in case["commits"]:
            model.commit(_c(commit))
        for read in case["reads"]:
            expected = _c(read["expected"])
            got = model.read(read["room_elements"])
            # sample counts are exact (FirTests.cpp:75,185-186 assert used() bytes)
            assert got.size == expected.size, case["name"]
            assert np.max(np.abs(got - expected)) < fir_kat["tolerance"], case["name"]


def test_fir_kat_counts_match_reference_rule(fir_kat):
    # 5 inputs, T=2, D=2 -> 2; 8 inputs, T=3, D=2 -> 3 (1 then 2)
    assert orc.fir_num_outputs_verbatim(5, 2, 2) == 2 == orc.fir_num_outputs(5, 2, 2)
    assert orc.fir_num_outputs_verbatim(8, 3, 2) == 3 == orc.fir_num_outputs(8, 3, 2)
    assert orc.fir_num_outputs_verbatim(6, 3, 2) == 2 == orc.fir_num_outputs(6, 3, 2)


def test_count_rule_equals_verbatim_where_reference_is_well_defined():
    # Fir.cpp:181-186 wraps when D >= T+2 or when T-D+1 <= nIn < T-1; everywhere else both agree.
    for T in range(1, 40):
        for D in range(1, T + 2):
            for n in range(0, 200):
                if n < T - 1 and n >= T - D + 1:
                    continue  # reference underflows nIn-(T-1)
                assert orc.fir_num_outputs(n, T, D) == orc.fir_num_outputs_verbatim(n, T, D), (n, T, D)


def test_c1_count_is_the_reference_one_fewer():
    # C1 2^20 inputs, T=63, D=10 -> 104851 (maths would allow 104852)
    assert orc.fir_num_outputs(1 << 20, 63, 10) == 104851
    assert orc.fir_num_outputs(4, 2, 2) == 1


def test_cosine_kat(cosine_kat):
    fs, f = float(cosine_kat["sample_rate"]), float(cosine_kat["frequency"])
    n_checked = cosine_kat["output_value_count"]
    align = cosine_kat["allocator_alignment"]
    n_emitted = (n_checked * 8 + align - 1) // align * align // 8
    assert n_emitted == 104
    delta = orc.lib().orc_cosine_delta(fs, f)
    phi_end = orc.lib().orc_cosine_phi_end(0.0, n_emitted, delta)
    got = orc.cosine_c(0.0, phi_end, n_emitted)[:n_checked]
    i = np.arange(n_checked, dtype=np.float32)
    theta = i * np.float32(f) / np.float32(fs) * np.float32(np.pi) * np.float32(2.0)
    assert np.max(np.abs(got.real - np.cos(theta))) < cosine_kat["tolerance"]
    assert np.max(np.abs(got.imag - np.sin(theta))) < cosine_kat["tolerance"]


def test_int8_scale_is_exact_and_bit_stable():
    x = np.arange(-128, 128, dtype=np.int8)
    y = orc.int8_to_norm_float(x)
    assert y.dtype == np.float32
    assert np.array_equal(y, x.astype(np.float32) / np.float32(128.0))
    assert y[0] == -1.0 and y[-1] == np.float32(127 / 128)


@pytest.mark.parametrize("chunks", [[5], [3, 2], [1, 1, 1, 1, 1], [2, 3]])
def test_stream_model_chunking_independent(chunks):
    rng = np.random.default_rng(7)
    taps = rng.standard_normal(7).astype(np.float32)
    x = (rng.standard_normal(64) + 1j * rng.standard_normal(64)).astype(np.complex64)
    whole = orc.fir("fc", taps, x, 3)
    model = orc.FirStreamModel("fc", taps, 3)
    outs, pos = [], 0
    while pos < x.size:
        for c in chunks:
            model.commit(x[pos:pos + c * 3])
            pos += c * 3
            outs.append(model.read())
            if pos >= x.size:
                break
    got = np.concatenate(outs)
    assert got.size == whole.size == (64 - 6) // 3
    assert np.allclose(got, whole, rtol=0, atol=1e-12)


def test_chain_matches_staged_ops():
    rng = np.random.default_rng(3)
    n = 4000
    iq = rng.integers(-128, 128, size=2 * n, dtype=np.int8)
    taps1 = rng.standard_normal(21).astype(np.float32)
    taps2 = rng.standard_normal(9).astype(np.float32)
    spec = orc.ChainSpec(1.0e6, -123456.0, taps1, 8, orc.FM, 0.7, taps2, 4)
    audio, rf, demod = orc.chain(spec, iq, n0=1000, want_rf=True, want_demod=True)
    # staged by hand in numpy fp64 with the exact phase definition
    x = (iq[0::2].astype(np.float64) + 1j * iq[1::2].astype(np.float64)) / 128.0
    idx = np.arange(n, dtype=np.uint64) + np.uint64(1000)
    turns = (idx * np.uint64(spec.step)).astype(np.int64).astype(np.float64) / 2.0**64
    z = x * np.exp(2j * np.pi * turns)
    n_rf = (n + 1 - 21) // 8
    y = np.array([np.dot(taps1.astype(np.float64), z[k * 8:k * 8 + 21]) for k in range(n_rf)])
    assert np.allclose(rf, y, rtol=0, atol=1e-12)
    d = float(np.float32(0.7)) * np.angle(y[1:] * np.conj(y[:-1]))
    assert np.allclose(demod, d, rtol=0, atol=1e-12)
    n_a = (d.size + 1 - 9) // 4
    a = np.array([np.dot(taps2.astype(np.float64), d[k * 4:k * 4 + 9]) for k in range(n_a)])
    assert audio.size == n_a and np.allclose(audio, a, rtol=0, atol=1e-12)


def test_phase_step_fixed_point():
    assert orc.phase_step(0.0, 1e6) == 0
    assert orc.phase_step(250e3, 1e6) == 1 << 62
    assert orc.phase_step(-250e3, 1e6) == (1 << 64) - (1 << 62)
    assert orc.phase_step(1.5e6, 1e6) == 1 << 63
