"""GPU parity of the wideband channelizer (BASELINE config C5) against the fp64 oracle run channel by channel, called
through the b200sdr C-ABI; shape and count edge cases; channel sharding equals the unsharded result bit for bit."""
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


def make(sdr, fs, freqs, mods, T1, D1, T2, D2, dev_hz):
    t1 = sdr.taps.lowpass(T1, 0.4 * fs / D1, fs)
    t2 = sdr.taps.lowpass(T2, 0.4 * fs / D1 / D2, fs / D1)
    gains = [sdr.fm_gain(fs / D1, dev_hz)] * len(freqs)
    return sdr.Channelizer(fs, freqs, mods, t1, D1, t2, D2, fm_gains=gains), t1, t2, gains


def check(sdr, fs, freqs, mods, T1, D1, T2, D2, n, dev_hz=5e3, seed=3):
    ch, t1, t2, gains = make(sdr, fs, freqs, mods, T1, D1, T2, D2, dev_hz)
    x = sdr.synth.int8_iq(n, seed=seed, sample_rate=fs)
    out, counts = ch.process(torch.from_numpy(x).to(DEV))
    got = out.cpu().numpy()
    n_demod, n_audio = ch.counts(n)
    assert min(counts) == n_audio and got.shape == (len(freqs), max(counts))
    for c, (f, m) in enumerate(zip(freqs, mods)):
        spec = orc.ChainSpec(fs, f, t1, D1, m, gains[c], t2, D2)
        ref, _, _ = orc.chain(spec, x)
        # counts are exact PER CHANNEL: an AM channel next to FM channels emits its own single-chain count (Fir.cpp:181-186
        # applied to that channel's demodulator output), not the FM channels' count
        assert counts[c] == ref.size == ch.channel_counts(c, n)[1], (c, m, counts[c], ref.size)
        tol = 2e-5 if m == 1 else 1e-5
        assert_close(got[c, :counts[c]], ref, tol=tol, what=f"channel {c} ({'FM' if m else 'AM'}, f={f})")
    return ch, got[:, :n_audio]


@pytest.mark.parametrize("tc", ["1", "0"])
def test_small_mixed_channels(sdr, monkeypatch, tc):
    monkeypatch.setenv("B200SDR_CHANNEL_TC", tc)
    fs = 1.024e6
    freqs = [-300e3, -111e3, 50e3, 222e3, 333e3, -7e3]
    mods = [0, 1, 0, 1, 1, 0]
    ch, _ = check(sdr, fs, freqs, mods, T1=400, D1=64, T2=33, D2=5, n=200003)
    assert ch.variant.startswith("channel<tcgen05" if tc == "1" else "channel<imma"), ch.variant


def test_few_taps_per_phase_uses_one_n_tile(sdr):
    fs = 2.0e6
    ch, _ = check(sdr, fs, [100e3, -400e3, 650e3], [0, 0, 0], T1=101, D1=40, T2=65, D2=10, n=150001)
    assert "NTC=1" in ch.variant


@pytest.mark.parametrize("tc", ["1", "0"])
def test_c5_shape_eight_channels(sdr, monkeypatch, tc):
    """C5 per-channel shape (4097 taps, decimate by 640, 273 audio taps, decimate by 5) on an 8-channel slice, off any raster:
    the per-channel int8 GEMM, on tcgen05 (UTCIMMA, accumulators in TMEM) and on the legacy IMMA path."""
    monkeypatch.setenv("B200SDR_CHANNEL_TC", tc)
    fs = 153.6e6
    freqs = [(-4 + i) * 600e3 + 37e3 for i in range(8)]
    mods = [i & 1 for i in range(8)]
    check(sdr, fs, freqs, mods, T1=4097, D1=640, T2=273, D2=5, n=(1 << 21) + 777, dev_hz=75e3)


@pytest.mark.parametrize("pfb", ["1", "0"])
def test_am_channels_keep_their_own_count_next_to_fm(sdr, monkeypatch, pfb):
    """Input lengths swept over one audio-decimation period of RF outputs: for exactly one residue the AM channels own one
    more audio sample than the FM channels (the FM discriminator holds one RF output back); both routes must emit it."""
    monkeypatch.setenv("B200SDR_PFB", pfb)
    fs, T1, D1, T2, D2 = 1.024e6, 400, 64, 33, 5
    freqs = [7e3 + b * fs / 16 for b in (0, 3, 5, 15, 9)]
    mods = [0, 1, 0, 1, 1]
    differ = 0
    for extra_rf in range(D2):
        n = T1 - 1 + D1 * (2000 + extra_rf) + 17
        ch, _ = check(sdr, fs, freqs, mods, T1, D1, T2, D2, n=n, seed=40 + extra_rf)
        am, fm_ = ch.channel_counts(0, n)[1], ch.channel_counts(1, n)[1]
        assert am - fm_ in (0, 1)
        differ += am - fm_
    assert differ == 1


def test_counts_and_short_inputs(sdr):
    fs = 1.024e6
    ch, *_ = make(sdr, fs, [10e3, 20e3], [0, 1], 400, 64, 33, 5, 5e3)
    for n in (0, 1, 399, 463, 400 + 64 * 40):
        n_demod, n_audio = ch.counts(n)
        assert n_audio == orc.chain_num_outputs(n, 400, 64, 1, 33, 5)
        x = torch.zeros(2 * max(n, 8), dtype=torch.int8, device=DEV)[: 2 * n]
        assert ch.run(x).shape == (2, n_audio)


def test_channel_sharding_is_bit_exact(sdr, monkeypatch):
    """Sharding by channel (the multi-GPU decomposition): each shard's channels equal the same channels of the full set.
    (Per-channel route: a shard that happens to fit a raster would take the filter bank, whose results agree to the parity
    tolerance, not bit for bit -- test_filter_bank_equals_direct_route_on_a_shard.)"""
    monkeypatch.setenv("B200SDR_PFB", "0")
    from cuda_sdr_b200 import sharding
    fs = 1.024e6
    freqs = [(-5 + i) * 90e3 + 1e3 for i in range(11)]
    mods = [i % 2 for i in range(11)]
    full, t1, t2, gains = make(sdr, fs, freqs, mods, 400, 64, 33, 5, 5e3)
    x = torch.from_numpy(sdr.synth.int8_iq(120001, seed=9, sample_rate=fs)).to(DEV)
    whole = full.run(x)
    n_audio = whole.shape[1]
    for world in (2, 4):
        for rank in range(world):
            mine = sharding.channels_of_rank(len(freqs), world, rank)
            part = sdr.Channelizer(fs, [freqs[i] for i in mine], [mods[i] for i in mine], t1, 64, t2, 5,
                                   fm_gains=[gains[i] for i in mine]).run(x, n_audio)
            assert torch.equal(part.view(torch.int32), whole[mine].view(torch.int32)), (world, rank)


# ---- polyphase-filter-bank route (pfb_kernels.cuh): channels on a raster fs/N with one common offset ----------------------
@pytest.mark.parametrize("pfb", ["1", "0"])
@pytest.mark.parametrize("log2n,bins,mods", [
    (4, [0, 3, 5, 15, 9, 8], [0, 1, 0, 1, 1, 0]),          # N = 16: two radix-4 passes
    (5, [1, 31, 16, 7], [0, 0, 0, 0]),                    # N = 32: radix-4, radix-4, radix-2; AM only (no successor output)
    (7, [0, 127, 64, 1, 100, 33, 77], [1, 1, 0, 1, 0, 1, 1]),  # N = 128
])
def test_raster_channels_take_the_filter_bank(sdr, monkeypatch, pfb, log2n, bins, mods):
    monkeypatch.setenv("B200SDR_PFB", pfb)
    fs, n_fft = 1.024e6, 1 << log2n
    freqs = [7e3 + b * fs / n_fft for b in bins]  # bins above N/2 are negative frequencies (turns are mod 1)
    ch, _ = check(sdr, fs, freqs, mods, T1=400, D1=64, T2=33, D2=5, n=200003)
    assert ch.variant.startswith(f"pfb<N={n_fft},fp64>" if pfb == "1" else "channel<"), ch.variant


@pytest.mark.parametrize("env", [{}, {"B200SDR_PFB256_GRID": "3"}, {"B200SDR_PFB256_GRID": "1"}, {"B200SDR_PFB256": "0"}])
def test_filter_bank_c5_shape_and_empty_channels(sdr, monkeypatch, env):
    """C5 shape on the 1/256 raster; most of the 24 channels carry no signal at all, so their level is set by what leaks
    from the strong ones: the bound is relative to each channel's own level.  Default kernel = pfb256 (taps in registers,
    register FFT, TMA input ring); with the grid limited to a few CTAs each one walks hundreds of rounds, so the input
    ring wraps many times and the lagging FM flush runs on every block; B200SDR_PFB256=0 is the first filter-bank kernel."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    fs = 153.6e6
    freqs = [(c - 12) * 600e3 + 100e3 for c in range(24)]
    mods = [c & 1 for c in range(24)]
    ch, _ = check(sdr, fs, freqs, mods, T1=4097, D1=640, T2=273, D2=5, n=(1 << 21) + 12345, dev_hz=75e3)
    assert ch.variant.startswith("pfb<N=256,fp64>"), ch.variant
    assert ("kernel=pfb256" in ch.variant) == (env.get("B200SDR_PFB256") != "0"), ch.variant


@pytest.mark.parametrize("mods", [[0] * 7, [1] * 7, [0, 1, 1, 0, 1, 0, 0]])
def test_pfb256_am_only_fm_only_and_counts(sdr, mods):
    """The N = 256 kernel without FM channels (no lag, flush at the end of a block), with FM channels only, and mixed sets over
    one audio-decimation period of input lengths (for one residue the AM channels own one more output: the AM tail pass)."""
    fs, T1, D1, T2, D2 = 153.6e6, 4097, 640, 273, 5
    freqs = [(c - 3) * 600e3 * 7 + 100e3 for c in range(7)]
    differ = 0
    for extra_rf in range(D2 if 0 in mods and 1 in mods else 1):
        n = T1 - 1 + D1 * (1500 + extra_rf) + 33
        ch, _ = check(sdr, fs, freqs, mods, T1, D1, T2, D2, n=n, dev_hz=75e3, seed=50 + extra_rf)
        assert "kernel=pfb256" in ch.variant, ch.variant
        counts = [ch.channel_counts(c, n)[1] for c in range(7)]
        differ += max(counts) - min(counts)
    assert differ == (1 if 0 in mods and 1 in mods else 0)


def test_pfb256_time_segments_concatenate_bit_exactly(sdr, monkeypatch):
    """Every RF output is a function of its own input window only: overlapped time segments (the multi-GPU decomposition
    of the filter-bank route) equal the one-shot result bit for bit, whatever the CTA partition of each launch."""
    fs, T1, D1, T2, D2 = 153.6e6, 4097, 640, 273, 5
    freqs = [(c - 6) * 600e3 * 3 + 100e3 for c in range(12)]
    mods = [c & 1 for c in range(12)]
    ch, t1, t2, gains = make(sdr, fs, freqs, mods, T1, D1, T2, D2, 75e3)
    assert "kernel=pfb256" in ch.variant, ch.variant
    n = (1 << 22) + 4097
    x = torch.from_numpy(sdr.synth.int8_iq(n, seed=22, sample_rate=fs)).to(DEV)
    whole = ch.run(x)
    n_audio = whole.shape[1]
    for parts in (2, 3, 8):
        cols = []
        for i in range(parts):
            a0, cnt, i0, icnt = ch.segment(n_audio, parts, i)
            cols.append(ch.run(x[2 * i0: 2 * (i0 + icnt)], cnt))
        cat = torch.cat(cols, dim=1)
        assert cat.shape == whole.shape
        assert torch.equal(cat.view(torch.int32), whole.view(torch.int32)), parts
    monkeypatch.setenv("B200SDR_PFB256_GRID", "2")  # another CTA partition of the same launch
    again = ch.run(x)
    assert torch.equal(again.view(torch.int32), whole.view(torch.int32))


def test_filter_bank_equals_direct_route_on_a_shard(sdr, monkeypatch):
    """A shard of the raster may resolve to a coarser raster (smaller N): results agree to the parity tolerance."""
    fs = 1.024e6
    freqs = [7e3 + b * fs / 64 for b in (0, 8, 16, 40)]  # this subset also fits N = 8
    mods = [0, 1, 0, 1]
    ch, got = check(sdr, fs, freqs, mods, T1=400, D1=64, T2=33, D2=5, n=100001)
    assert ch.variant.startswith("pfb<N=8,fp64>"), ch.variant
    monkeypatch.setenv("B200SDR_PFB", "0")
    _, ref = check(sdr, fs, freqs, mods, T1=400, D1=64, T2=33, D2=5, n=100001)
    for c in range(4):
        assert_close(got[c], ref[c], tol=2e-5 if mods[c] else 1e-5, what=f"channel {c}")


def test_filter_bank_time_segments_concatenate_bit_exactly(sdr):
    """The multi-GPU decomposition of the filter-bank route: overlapped time segments of the wideband stream, all channels
    per segment.  Every RF output is a function of its own input window only, so segments equal the one-shot result."""
    from cuda_sdr_b200 import sharding
    fs, T1, D1, T2, D2 = 1.024e6, 400, 64, 33, 5
    bins = [0, 3, 5, 15, 9, 8, 12]
    mods = [0, 1, 0, 1, 1, 0, 1]
    freqs = [7e3 + b * fs / 16 for b in bins]
    ch, t1, t2, gains = make(sdr, fs, freqs, mods, T1, D1, T2, D2, 5e3)
    assert ch.variant.startswith("pfb<N=16")
    n = 300007
    x = torch.from_numpy(sdr.synth.int8_iq(n, seed=21, sample_rate=fs)).to(DEV)
    whole = ch.run(x)
    n_audio = whole.shape[1]
    window = sharding.chain_window(T1, D1, T2, True, True)
    for parts in (2, 3, 8):
        cols = []
        for i in range(parts):
            seg = sharding.time_segment(n_audio, parts, i, D1 * D2, window)
            cols.append(ch.run(x[2 * seg.first_input: 2 * (seg.first_input + seg.input_count)], seg.output_count))
        cat = torch.cat(cols, dim=1)
        assert cat.shape == whole.shape
        assert torch.equal(cat.view(torch.int32), whole.view(torch.int32)), parts


# ---- the BASELINE C5 shape at its full channel count ----------------------------------------------------------------------
_C5_CACHE = {}


@pytest.mark.parametrize("pfb", ["256", "1", "0", "0-imma"])
def test_c5_256_channels_against_the_oracle(sdr, monkeypatch, pfb):
    """BASELINE configs[4] as the bench runs it: 256 channels on the 600 kHz raster alternating AM/FM, 4097 taps / 640, 273
    audio taps / 5, on 2^24 samples -- EVERY channel of all four routes (the N = 256 filter-bank kernel the bench runs, the
    first filter-bank kernel, the per-channel int8 GEMM on tcgen05 / TMEM and on the legacy IMMA path) against the fp64 oracle run channel by channel, with per-channel exact counts."""
    monkeypatch.setenv("B200SDR_PFB", "0" if pfb.startswith("0") else "1")
    monkeypatch.setenv("B200SDR_PFB256", "1" if pfb == "256" else "0")
    monkeypatch.setenv("B200SDR_CHANNEL_TC", "0" if pfb == "0-imma" else "1")
    fs, T1, D1, T2, D2 = 153.6e6, 4097, 640, 273, 5
    total = 256
    freqs = [(c - total / 2) * 600e3 + 100e3 for c in range(total)]
    mods = [c & 1 for c in range(total)]
    t1 = sdr.taps.lowpass(T1, 100e3, fs)
    t2 = sdr.taps.lowpass(T2, 0.45 * 48e3, fs / D1)
    gain = sdr.fm_gain(fs / D1, 75e3)
    ch = sdr.Channelizer(fs, freqs, mods, t1, D1, t2, D2, fm_gains=[gain] * total)
    expect = {"256": "pfb<N=256,fp64>", "1": "pfb<N=256,fp64>", "0": "channel<tcgen05", "0-imma": "channel<imma"}[pfb]
    assert ch.variant.startswith(expect), ch.variant
    assert ("kernel=pfb256" in ch.variant) == (pfb == "256")
    n = (1 << 24) + 640 * 3 + 11
    if "x" not in _C5_CACHE:  # the input and the 256 oracle runs are shared by the two routes
        _C5_CACHE["x"] = sdr.synth.int8_iq(n, seed=0x5D120005 & 0xffff, sample_rate=fs)
        _C5_CACHE["ref"] = [orc.chain(orc.ChainSpec(fs, freqs[c], t1, D1, mods[c], gain, t2, D2), _C5_CACHE["x"])[0] for c in range(total)]
    x = _C5_CACHE["x"]
    out, counts = ch.process(torch.from_numpy(x).to(DEV))
    got = out.cpu().numpy()
    for c in range(total):
        ref = _C5_CACHE["ref"][c]
        assert counts[c] == ref.size, (c, counts[c], ref.size)
        tol = 2e-5 if mods[c] else 1e-5
        assert_close(got[c, :counts[c]], ref, tol=tol, what=f"channel {c} ({'FM' if mods[c] else 'AM'})")
