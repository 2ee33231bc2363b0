"""GPU parity of the chain -- the fused persistent kernels (toepKernel: RF stage as an int8 GEMM over a Toeplitz view of
the input; chainKernel: polyphase rows) and the two-kernel path (K1 =
convert+mix+FIR+demod, K2 = audio FIR) -- called through the b200sdr C-ABI, against the fp64 CPU oracle; plus
size-independent properties at BASELINE sizes."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.util import assert_close, assert_fm_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def sdr():
    import cuda_sdr_b200 as m
    return m


def c2_spec(sdr, fs=19.2e6):
    """BASELINE configs[1] (AM): mix -> 101-tap D=40 -> |.| -> 129-tap D=10 -> 48 kHz."""
    t1 = sdr.taps.lowpass(101, 0.45 * fs / 40, fs)
    t2 = sdr.taps.lowpass(129, 0.45 * 48e3, fs / 40)
    return dict(sample_rate=fs, frequency=-1.234e6, rf_taps=t1, rf_decim=40, modulation=sdr.AM, audio_taps=t2, audio_decim=10)


def c3_spec(sdr, fs=19.2e6):
    """BASELINE configs[2] (WBFM): mix -> 545-tap D=80 -> FM discriminator -> 273-tap D=5 -> 48 kHz."""
    t1 = sdr.taps.lowpass(545, 100e3, fs)
    t2 = sdr.taps.lowpass(273, 0.45 * 48e3, fs / 80)
    g = sdr.fm_gain(fs / 80, 75e3)
    return dict(sample_rate=fs, frequency=2.5e6, rf_taps=t1, rf_decim=80, modulation=sdr.FM, fm_gain=g, audio_taps=t2, audio_decim=5)


def oracle_spec(kw, input_int8=True, mix=True):
    return orc.ChainSpec(kw["sample_rate"], kw["frequency"], kw["rf_taps"], kw["rf_decim"], kw.get("modulation", 0),
                         kw.get("fm_gain", 1.0), kw.get("audio_taps"), kw.get("audio_decim", 1), mix=mix, input_int8=input_int8)


def check_chain(sdr, kw, x_host, n0=0, input_int8=True, mix=True, what=""):
    chain = sdr.Chain(**kw, input_int8=input_int8, mix=mix)
    spec = oracle_spec(kw, input_int8, mix)
    dx = torch.from_numpy(x_host).to(DEV)
    n_in = x_host.size // 2 if input_int8 else x_host.size
    n_rf, n_demod, n_audio = chain.counts(n_in)
    audio_ref, rf_ref, demod_ref = orc.chain(spec, x_host, n0=n0, want_demod=True)
    # stage K1 alone (demodulated samples)
    if kw.get("modulation", 0) != sdr.NONE:
        demod = chain.rf_stage(dx, n_demod, n0).cpu().numpy()
        assert demod.size == demod_ref.size
        if kw.get("modulation", 0) == sdr.FM:
            assert_fm_close(demod, demod_ref, kw["fm_gain"], what=what + " K1")
        else:
            assert_close(demod, demod_ref, what=what + " K1")
    got = chain.process_device(dx, n0).cpu().numpy()
    assert got.size == audio_ref.size == n_audio, (got.size, audio_ref.size, n_audio)
    assert_close(got, audio_ref, tol=2e-5 if kw.get("modulation", 0) == sdr.FM else 1e-5, what=what + " chain")
    return chain, got


@pytest.mark.parametrize("n", [1 << 20, 777777])
def test_c2_am_chain(sdr, n):
    kw = c2_spec(sdr)
    x = sdr.synth.int8_iq(n)
    chain, _ = check_chain(sdr, kw, x, n0=12345, what="C2")
    assert chain.variant.startswith("toeplitz<int8c,G=2"), chain.variant


@pytest.mark.parametrize("n", [1 << 20, 654321])
def test_c3_wbfm_chain(sdr, n):
    kw = c3_spec(sdr)
    x = sdr.synth.int8_iq(n)
    chain, _ = check_chain(sdr, kw, x, n0=99, what="C3")
    assert chain.variant.startswith(("toeplitz<int8c", "rows<int8c,mix=1,MP=7")), chain.variant


def test_none_mode_outputs_mixed_rf_samples_with_absolute_phase(sdr):
    kw = c2_spec(sdr)
    kw.update(modulation=sdr.NONE, audio_taps=None)
    x = sdr.synth.int8_iq(300000)
    chain = sdr.Chain(**kw)
    spec = oracle_spec(kw)
    for n0 in (0, 1, 1 << 33):
        got = chain.process_device(torch.from_numpy(x).to(DEV), n0).cpu().numpy()
        ref, _, _ = orc.chain(spec, x, n0=n0)
        assert got.size == ref.size
        assert_close(got, ref, what=f"NONE n0={n0}")


@pytest.mark.parametrize("mix", [True, False])
@pytest.mark.parametrize("mod", [0, 1])
def test_cf32_input_and_no_mixer(sdr, mix, mod):
    fs = 2.4e6
    t1 = sdr.taps.lowpass(65, 90e3, fs)
    t2 = sdr.taps.lowpass(33, 20e3, fs / 10)
    kw = dict(sample_rate=fs, frequency=300e3, rf_taps=t1, rf_decim=10, modulation=mod, fm_gain=0.9, audio_taps=t2, audio_decim=5)
    x = sdr.synth.cf32(200003, seed=mod + 2 * mix)
    check_chain(sdr, kw, x, n0=5, input_int8=False, mix=mix, what=f"cf32 mix={mix} mod={mod}")


@pytest.mark.parametrize("toeplitz", ["1", "0"])
@pytest.mark.parametrize("T1,D1", [(101, 40), (40, 40), (17, 64), (81, 24), (33, 7), (640, 16), (301, 100), (3, 8), (1000, 8)])
def test_shapes_cover_rows_and_direct_paths(sdr, T1, D1, monkeypatch, toeplitz):
    monkeypatch.setenv("B200SDR_TOEPLITZ", toeplitz)
    fs = 1.0e6
    kw = dict(sample_rate=fs, frequency=-77e3, rf_taps=sdr.taps.lowpass(T1, 0.4 * fs / D1, fs), rf_decim=D1, modulation=sdr.AM,
              audio_taps=sdr.taps.lowpass(21, 0.2 * fs / D1, fs / D1), audio_decim=3)
    x = sdr.synth.int8_iq(150001, seed=T1 * 7 + D1)
    check_chain(sdr, kw, x, what=f"T1={T1} D1={D1}")


def test_empty_and_short_inputs(sdr):
    kw = c2_spec(sdr)
    chain = sdr.Chain(**kw)
    for n in (0, 1, 100, 101 + 39, 101 + 40 * 128):  # < T1, first RF output, still < T2 demod samples
        x = torch.zeros(2 * n, dtype=torch.int8, device=DEV)
        out = chain.process_device(x)
        assert out.numel() == orc.chain_num_outputs(n, 101, 40, 0, 129, 10)
    # first audio output: 129 demod samples + (D2-1) for the count rule, each RF output rule likewise
    n1 = 101 - 1 + 40 * (129 - 1 + 10)
    assert chain.counts(n1)[2] == 1 == orc.chain_num_outputs(n1, 101, 40, 0, 129, 10) and chain.counts(n1 - 1)[2] == 0


@pytest.mark.parametrize("toeplitz", ["1", "0"])
def test_time_segments_concatenate_bit_exactly(sdr, monkeypatch, toeplitz):
    monkeypatch.setenv("B200SDR_FUSED", "1")
    monkeypatch.setenv("B200SDR_TOEPLITZ", toeplitz)
    """Outputs are a pure function of the absolute sample index: overlapped time segments (the
    multi-GPU and host-staging decomposition) must reproduce the one-shot result bit for bit."""
    kw = c3_spec(sdr)
    chain = sdr.Chain(**kw)
    n = 1 << 20
    x = torch.from_numpy(sdr.synth.int8_iq(n)).to(DEV)
    whole = chain.process_device(x)
    n_audio = whole.numel()
    assert chain.fused and chain.variant.startswith("toeplitz<" if toeplitz == "1" else "chain<")
    for parts in (2, 3, 8):
        outs = []
        for i in range(parts):
            a0, cnt, i0, icnt = chain.segment(n_audio, parts, i)
            seg = x[2 * i0: 2 * (i0 + icnt)]
            outs.append(chain.run(seg, cnt, i0, n_in=icnt))
        cat = torch.cat(outs)
        assert cat.numel() == n_audio
        assert torch.equal(cat.view(torch.int32), whole.view(torch.int32)), f"parts={parts}"


def test_two_kernel_path_segments_concatenate_bit_exactly(sdr, monkeypatch):
    monkeypatch.setenv("B200SDR_FUSED", "0")
    kw = c3_spec(sdr)
    chain = sdr.Chain(**kw)
    assert not chain.fused
    n = 1 << 19
    x = torch.from_numpy(sdr.synth.int8_iq(n)).to(DEV)
    whole = chain.process_device(x)
    n_audio = whole.numel()
    outs = []
    for i in range(3):
        a0, cnt, i0, icnt = chain.segment(n_audio, 3, i)
        seg = x[2 * i0: 2 * (i0 + icnt)]
        demod = chain.rf_stage(seg, (cnt - 1) * chain.audio_decim + chain.T2, i0, n_in=icnt)
        outs.append(chain.audio_stage(demod, cnt))
    assert torch.equal(torch.cat(outs).view(torch.int32), whole.view(torch.int32))


@pytest.mark.parametrize("env", [
    {"B200SDR_FUSED": "0"}, {"B200SDR_FUSED": "1"},
    {"B200SDR_FUSED": "1", "B200SDR_CHAIN_RPT": "2"}, {"B200SDR_FUSED": "1", "B200SDR_CHAIN_RPT": "4"},
    {"B200SDR_FUSED": "1", "B200SDR_CHAIN_STAGES": "1"}, {"B200SDR_FUSED": "1", "B200SDR_CHAIN_STAGES": "2", "B200SDR_CHAIN_CTAS": "1"},
    {"B200SDR_FUSED": "1", "B200SDR_CHAIN_MMA": "0"}, {"B200SDR_FUSED": "1", "B200SDR_CHAIN_MMA": "1"},
    {"B200SDR_FUSED": "1", "B200SDR_CHAIN_WARPS": "2"}, {"B200SDR_FUSED": "1", "B200SDR_CHAIN_WARPS": "8"},
    {"B200SDR_FUSED": "1", "B200SDR_CHAIN_RPT": "1", "B200SDR_CHAIN_WARPS": "12"}, {"B200SDR_FUSED": "1", "B200SDR_CHAIN_AUDIO_WARPS": "3"},
    {"B200SDR_FUSED": "1", "B200SDR_CHAIN_MMA": "0", "B200SDR_CHAIN_RPT": "1"},
    {"B200SDR_FUSED": "1", "B200SDR_CHAIN_STAGES": "3", "B200SDR_CHAIN_RPT": "2"},
    {"B200SDR_TOEPLITZ": "1"}, {"B200SDR_TOEPLITZ": "1", "B200SDR_TOEP_G": "1"}, {"B200SDR_TOEPLITZ": "1", "B200SDR_TOEP_STAGES": "3"},
    {"B200SDR_TOEPLITZ": "1", "B200SDR_TOEP_WARPS": "3"}, {"B200SDR_TOEPLITZ": "1", "B200SDR_TOEP_WARPS": "8", "B200SDR_TOEP_G": "1"},
    {"B200SDR_TOEPLITZ": "1", "B200SDR_TOEP_MAGIC": "0"}, {"B200SDR_TOEPLITZ": "1", "B200SDR_TOEP_WARPS": "6", "B200SDR_TOEP_G": "1"},
    {"B200SDR_TOEPLITZ": "1", "B200SDR_TOEP_CTAS": "1", "B200SDR_TOEP_WARPS": "3"},
])
@pytest.mark.parametrize("which", ["c2", "c3"])
def test_every_kernel_variant_matches_the_oracle(sdr, monkeypatch, env, which):
    """Tile shape, ring depth, conversion route and audio split are tuning knobs: each must give the same
    results (to the north-star tolerance) and the same counts."""
    monkeypatch.setenv("B200SDR_TOEPLITZ", "0")  # the B200SDR_CHAIN_* knobs belong to chainKernel
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    kw = c2_spec(sdr) if which == "c2" else c3_spec(sdr)
    x = sdr.synth.int8_iq(400000 + 17, seed=11)
    chain, _ = check_chain(sdr, kw, x, n0=4242, what=f"{which} {env}")
    if env.get("B200SDR_TOEPLITZ") == "1":
        assert chain.variant.startswith("toeplitz<")
    else:
        assert chain.fused == (env.get("B200SDR_FUSED") == "1") and not chain.variant.startswith("toeplitz<")


def test_default_routes(sdr, monkeypatch):
    assert sdr.Chain(**c2_spec(sdr)).variant.startswith("toeplitz<int8c,G=2,magic>")
    assert sdr.Chain(**c3_spec(sdr)).variant.startswith("toeplitz<int8c,G=2,i2f>")
    monkeypatch.setenv("B200SDR_TOEPLITZ", "0")
    assert sdr.Chain(**c2_spec(sdr)).variant.startswith("chain<int8c,mix=1,MP=3,RPT=2,conv=imma>")
    assert sdr.Chain(**c3_spec(sdr)).variant.startswith("rows<int8c,mix=1,MP=7")


def test_host_path_equals_device_path(sdr):
    kw = c2_spec(sdr)
    chain = sdr.Chain(**kw)
    n = (1 << 21) + 12345
    xh = torch.from_numpy(sdr.synth.int8_iq(n)).pin_memory()
    dev = chain.process_device(xh.to(DEV)).cpu()
    chain.set_host_segment(1 << 18)  # force several staged segments
    host = chain.process_host(xh)
    assert host.numel() == dev.numel()
    assert torch.equal(host.view(torch.int32), dev.view(torch.int32))


def test_c2_full_size_properties(sdr):
    """BASELINE size (2^28 int8 IQ samples): exact output count, and random windows of the full-size
    result checked against the oracle run on just the input slice each window depends on."""
    kw = c2_spec(sdr)
    chain = sdr.Chain(**kw)
    n = 1 << 28
    x = sdr.synth.device_int8_iq(n, DEV)
    out = chain.process_device(x)
    assert out.numel() == orc.chain_num_outputs(n, 101, 40, 0, 129, 10) == chain.counts(n)[2]
    assert bool(torch.isfinite(out).all())
    spec = oracle_spec(kw)
    rng = np.random.default_rng(1)
    for a0 in [0, out.numel() - 64] + list(rng.integers(0, out.numel() - 64, size=4)):
        a0 = int(a0)
        i0 = a0 * chain.stride
        icnt = 63 * chain.stride + chain.window + chain.stride + chain.rf_decim  # the count rules need D-1 extra per stage
        icnt = min(icnt, n - i0)
        xs = x[2 * i0: 2 * (i0 + icnt)].cpu().numpy()
        ref, _, _ = orc.chain(spec, xs, n0=i0)
        m = min(ref.size, 64)
        assert_close(out[a0:a0 + m].cpu().numpy(), ref[:m], what=f"window@{a0}")


@pytest.mark.parametrize("env", [
    {"B200SDR_TOEP_GRID": "1"}, {"B200SDR_TOEP_GRID": "3"}, {"B200SDR_TOEP_GRID": "7"},
    {"B200SDR_TOEP_GRID": "2", "B200SDR_TOEP_G": "1"}, {"B200SDR_TOEP_GRID": "5", "B200SDR_TOEP_WARPS": "3"},
    {"B200SDR_TOEP_GRID": "2", "B200SDR_TOEP_AUDIO_WARPS": "1"}, {"B200SDR_TOEP_GRID": "4", "B200SDR_TOEP_AUDIO_WARPS": "3"},
    {"B200SDR_TOEP_GRID": "3", "B200SDR_TOEP_STAGES": "3"},
])
@pytest.mark.parametrize("which", ["c2", "c3"])
def test_demod_ring_wraps_many_times(sdr, monkeypatch, env, which):
    """toepKernel's demod ring holds 4 tiles and wraps only when a CTA walks more than that.  With the grid limited to a
    few CTAs every CTA runs >= 9 tiles on a 2^22-sample input, so the ring (and its mirror past the end) wraps at least
    twice: AM with even D2 = 10 (paired loads) and FM with OTW = 64G - 1, odd D2 = 5 (scalar loads), the shape the bench
    runs for C3."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    kw = c2_spec(sdr) if which == "c2" else c3_spec(sdr)
    n = (1 << 22) + 4321
    x = sdr.synth.int8_iq(n, seed=23)
    chain, _ = check_chain(sdr, kw, x, n0=777, what=f"{which} {env}")
    assert chain.variant.startswith("toeplitz<"), chain.variant
    grid = int(env["B200SDR_TOEP_GRID"])
    assert f"grid={grid})" in chain.variant, chain.variant
    # tiles per CTA: demod samples / (warps * (64 G - fm)) / grid
    import re
    m = re.search(r"G=(\d).*warps=(\d+)\+", chain.variant)
    ot = int(m.group(2)) * (64 * int(m.group(1)) - (1 if which == "c3" else 0))
    assert chain.counts(n)[1] / ot / grid >= 9.0


def test_c3_full_size_properties(sdr):
    """BASELINE size for the WBFM chain (2^28 int8 IQ samples, the shape `bench.py --workload wbfm` times): exact output
    count, finiteness, and random windows of the full-size result against the oracle run on just the input slice each
    window depends on.  Every CTA walks ~44 tiles here, so the FM demod ring wraps ~11 times."""
    kw = c3_spec(sdr)
    chain = sdr.Chain(**kw)
    assert chain.variant.startswith("toeplitz<"), chain.variant
    n = 1 << 28
    x = sdr.synth.device_int8_iq(n, DEV)
    out = chain.process_device(x)
    assert out.numel() == orc.chain_num_outputs(n, 545, 80, 1, 273, 5) == chain.counts(n)[2]
    assert bool(torch.isfinite(out).all())
    spec = oracle_spec(kw)
    rng = np.random.default_rng(2)
    for a0 in [0, out.numel() - 64] + list(rng.integers(0, out.numel() - 64, size=6)):
        a0 = int(a0)
        i0 = a0 * chain.stride
        icnt = 63 * chain.stride + chain.window + chain.stride + chain.rf_decim
        icnt = min(icnt, n - i0)
        xs = x[2 * i0: 2 * (i0 + icnt)].cpu().numpy()
        ref, _, _ = orc.chain(spec, xs, n0=i0)
        m = min(ref.size, 64)
        assert_close(out[a0:a0 + m].cpu().numpy(), ref[:m], tol=2e-5, what=f"window@{a0}")
    # segments of the full-size block concatenate bit-exactly (the multi-GPU decomposition at bench size)
    n_audio = out.numel()
    a0, cnt, i0, icnt = chain.segment(n_audio, 8, 5)
    part = chain.run(x[2 * i0: 2 * (i0 + icnt)], cnt, i0, n_in=icnt)
    assert torch.equal(part.view(torch.int32), out[a0:a0 + cnt].view(torch.int32))
