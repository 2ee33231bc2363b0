"""Host logic of the multi-GPU partitioning, incl. a world_size-2 gloo run of the gather (CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cuda_sdr_b200 import sharding
from oracle import oracle as orc


def test_segments_tile_the_outputs_and_cover_exactly_the_needed_inputs():
    T1, D1, T2, D2 = 101, 40, 129, 10
    for fm in (False, True):
        stride = sharding.chain_stride(D1, D2, True)
        window = sharding.chain_window(T1, D1, T2, fm, True)
        assert window == (T2 - 1 + fm) * D1 + T1
        n_audio = 1000
        for parts in (1, 2, 3, 8):
            segs = [sharding.time_segment(n_audio, parts, i, stride, window) for i in range(parts)]
            assert segs[0].first_output == 0 and sum(s.output_count for s in segs) == n_audio
            for a, b in zip(segs, segs[1:]):
                assert a.first_output + a.output_count == b.first_output
                # look-ahead halo: the next segment starts before this one ends
                assert b.first_input < a.first_input + a.input_count
            last = segs[-1]
            assert last.first_input + last.input_count == (n_audio - 1) * stride + window


def test_segmented_oracle_equals_whole_oracle():
    rng = np.random.default_rng(0)
    T1, D1, T2, D2 = 31, 8, 17, 4
    taps1 = rng.standard_normal(T1).astype(np.float32)
    taps2 = rng.standard_normal(T2).astype(np.float32)
    n = 20000
    iq = rng.integers(-128, 128, size=2 * n, dtype=np.int8)
    for mod in (orc.AM, orc.FM):
        spec = orc.ChainSpec(1e6, 12345.0, taps1, D1, mod, 0.5, taps2, D2)
        whole, _, _ = orc.chain(spec, iq)
        stride = sharding.chain_stride(D1, D2, True)
        window = sharding.chain_window(T1, D1, T2, mod == orc.FM, True)
        parts = 4
        outs = []
        for i in range(parts):
            s = sharding.time_segment(whole.size, parts, i, stride, window)
            # the reference count rule needs D-1 extra samples per stage to release a block's last output
            seg = iq[2 * s.first_input: 2 * (s.first_input + s.input_count + D1 * D2 + D1)]
            o, _, _ = orc.chain(spec, seg, n0=s.first_input)
            outs.append(o[: s.output_count])
        assert np.allclose(np.concatenate(outs), whole, rtol=0, atol=1e-12)


def test_channels_of_rank_is_a_partition():
    for world in (1, 2, 4, 8):
        got = sorted(c for r in range(world) for c in sharding.channels_of_rank(256, world, r))
        assert got == list(range(256))
        assert all(len(sharding.channels_of_rank(256, world, r)) == 256 // world for r in range(world))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_audio = 1001
        counts = [sharding.time_segment(n_audio, world, r, 400, 5261).output_count for r in range(world)]
        first = sharding.time_segment(n_audio, world, rank, 400, 5261).first_output
        local = torch.arange(first, first + counts[rank], dtype=torch.float32)
        out = sharding.gather_to_rank0(local, counts)
        if rank == 0:
            q.put(bool(torch.equal(out, torch.arange(n_audio, dtype=torch.float32))))
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_to_rank0_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=100)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok
