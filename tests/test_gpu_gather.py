"""b200sdr_gather (include/b200sdr/b200sdr.h): the one exchange step of the multi-GPU decomposition -- every rank's decimated audio
to rank 0 over NCCL on a side stream, slab by slab -- and the time-segment arithmetic of the channelizer.  A single rank needs no
NCCL; with two or more GPUs in the box two ranks run as two host threads of this process (one per device), which is the
single-process form of the same decomposition, and the sharded chain must equal the one-device result bit for bit."""
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sdr():
    import cuda_sdr_b200 as m
    return m


def test_single_rank_slab_ring(sdr):
    from cuda_sdr_b200 import sharding
    g = sharding.Gather(0, 1, [1000], slabs=3, device=0)
    dev = torch.device("cuda", 0)
    for k in range(7):  # more submissions than slabs: acquire() orders the reuse
        slab = k % 3
        g.acquire(slab)
        g.slab(slab).copy_(torch.arange(1000, dtype=torch.float32, device=dev) + k)
        g.submit(slab, [1000 - k])
        g.finish()
        torch.cuda.synchronize()
        got = g.result(slab, 0)[: 1000 - k].cpu().numpy()
        assert np.array_equal(got, np.arange(1000 - k, dtype=np.float32) + k)
    assert g.stats()["gathers"] == 7
    with pytest.raises(sdr._native.NativeError):
        g.submit(0, [1001])  # more than the slab holds
    g.close()


def test_channelizer_segments_tile_the_outputs(sdr):
    from cuda_sdr_b200 import sharding
    fs, T1, D1, T2, D2 = 1.024e6, 400, 64, 33, 5
    freqs = [7e3 + b * fs / 16 for b in (0, 3, 5)]
    for mods in ([0, 1, 0], [0, 0, 0]):
        ch = sdr.Channelizer(fs, freqs, mods, sdr.taps.lowpass(T1, 4e3, fs), D1, sdr.taps.lowpass(T2, 1e3, fs / D1), D2, fm_gains=[1.0] * 3)
        window = sharding.chain_window(T1, D1, T2, any(mods), True)
        for parts in (1, 2, 8):
            for i in range(parts):
                want = sharding.time_segment(977, parts, i, D1 * D2, window)
                assert ch.segment(977, parts, i) == (want.first_output, want.output_count, want.first_input, want.input_count)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in the box (gpurun --gpus 2)")
def test_two_ranks_shard_the_chain_bit_exactly(sdr):
    """Two ranks = two host threads, one per device: each runs its time segment of one stream (b200sdr_chain_segment), the audio is
    gathered to rank 0 by b200sdr_gather over NCCL; the result equals the one-device run bit for bit."""
    from cuda_sdr_b200 import sharding
    fs = 19.2e6
    t1 = sdr.taps.lowpass(101, 0.45 * fs / 40, fs)
    t2 = sdr.taps.lowpass(129, 0.45 * 48e3, fs / 40)
    n = (1 << 22) + 999
    x_host = sdr.synth.int8_iq(n, seed=77)
    whole_chain = sdr.Chain(fs, -1.234e6, t1, 40, sdr.AM, audio_taps=t2, audio_decim=10, device=0)
    whole = whole_chain.process_device(torch.from_numpy(x_host).to("cuda:0")).cpu()
    n_audio = whole.numel()
    world = 2
    segs = [whole_chain.segment(n_audio, world, r) for r in range(world)]
    uid = sharding.Gather.unique_id()
    results, errors = {}, []

    def rank_main(rank):
        try:
            torch.cuda.set_device(rank)
            dev = torch.device("cuda", rank)
            chain = sdr.Chain(fs, -1.234e6, t1, 40, sdr.AM, audio_taps=t2, audio_decim=10, device=rank)
            g = sharding.Gather(rank, world, [s[1] for s in segs], slabs=2, device=rank, unique_id=uid)
            a0, cnt, i0, icnt = segs[rank]
            seg = torch.from_numpy(x_host[2 * i0: 2 * (i0 + icnt)]).to(dev)
            for rep in range(3):  # the slab ring is reused
                slab = rep % 2
                g.acquire(slab)
                chain.run(seg, cnt, i0, out=g.slab(slab), n_in=icnt)
                g.submit(slab)
            g.finish()
            torch.cuda.synchronize(dev)
            if rank == 0:
                results["audio"] = torch.cat([g.result(0, r)[: segs[r][1]].cpu() for r in range(world)])
                results["nccl"] = g.stats()["nccl_version"]
            g.close()
        except Exception as e:  # surfaced by the main thread
            errors.append((rank, repr(e)))

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(300)
    assert not errors, errors
    assert results["nccl"] >= 20000
    assert torch.equal(results["audio"].view(torch.int32), whole.view(torch.int32))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in the box (gpurun --gpus 2)")
@pytest.mark.parametrize("transport", ["store", "copy"])
def test_peer_mode_gather_two_processes(transport):
    """The peer modes need one process per GPU (CUDA IPC): two ranks under torchrun, weighted (unequal) time segments, the chain
    kernel storing straight into rank 0's slabs over NVLink ("store") or into a local slab moved there by a copy engine ("copy"),
    seven rounds over two slabs; the gathered audio equals the one-device run bit for bit (tests/gather_peer_worker.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(root, "tests", "gather_peer_worker.py"), transport]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert res.returncode == 0 and "PEER_GATHER_OK" in res.stdout, (res.stdout + res.stderr)[-3000:]


def test_weighted_segments_tile_the_outputs(sdr):
    fs = 19.2e6
    chain = sdr.Chain(fs, -1.234e6, sdr.taps.lowpass(101, 0.45 * fs / 40, fs), 40, sdr.FM, fm_gain=0.1,
                      audio_taps=sdr.taps.lowpass(129, 0.45 * 48e3, fs / 40), audio_decim=10)
    for weights in ([1.0, 1.0, 1.0], [0.97, 1.03, 1.0, 1.01, 0.99, 1.0, 1.02, 0.98], [5.0, 1.0]):
        segs = [chain.segment_weighted(100003, weights, i) for i in range(len(weights))]
        assert segs[0][0] == 0 and segs[-1][0] + segs[-1][1] == 100003
        for a, b in zip(segs, segs[1:]):
            assert a[0] + a[1] == b[0]
        for (first, count, i0, icnt), w in zip(segs, weights):
            assert i0 == first * chain.stride and icnt == (count - 1) * chain.stride + chain.window
            assert abs(count - 100003 * w / sum(weights)) <= 1.0
    equal = [chain.segment_weighted(977, [1.0] * 4, i)[1] for i in range(4)]
    assert sum(equal) == 977 and max(equal) - min(equal) <= 1
