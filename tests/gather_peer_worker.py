"""Worker of tests/test_gpu_gather.py::test_peer_mode_gather_two_processes (run under torchrun, one process per GPU): every rank
runs its time segment of one stream with the chain kernel STORING INTO RANK 0'S MEMORY (b200sdr_gather, peer mode: CUDA IPC over
NVLink, stream-ordered flags) or, with argument "copy", into a local slab that a copy engine moves there; rank 0 checks the
gathered audio against its own single-device run, bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_sdr_b200 as sdr  # noqa: E402
from cuda_sdr_b200 import sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo")
    fs = 19.2e6
    t1 = sdr.taps.lowpass(101, 0.45 * fs / 40, fs)
    t2 = sdr.taps.lowpass(129, 0.45 * 48e3, fs / 40)
    n = (1 << 22) + 999
    x_host = sdr.synth.int8_iq(n, seed=77)
    chain = sdr.Chain(fs, -1.234e6, t1, 40, sdr.AM, audio_taps=t2, audio_decim=10, device=local)
    n_audio = chain.counts(n)[2]
    weights = [1.0 + 0.07 * r for r in range(world)]  # unequal shares: the balanced partition
    segs = [chain.segment_weighted(n_audio, weights, r) for r in range(world)]
    assert segs[0][0] == 0 and sum(s[1] for s in segs) == n_audio and all(segs[i][0] + segs[i][1] == segs[i + 1][0] for i in range(world - 1))
    g = sharding.Gather(rank, world, [s[1] for s in segs], slabs=2, device=local, mode=sharding.Gather.PEER_COPY if sys.argv[1:] == ["copy"] else sharding.Gather.PEER)
    blobs = [None] * world
    dist.all_gather_object(blobs, g.export_blob())
    g.import_blobs(b"".join(blobs))
    a0, cnt, i0, icnt = segs[rank]
    seg = torch.from_numpy(x_host[2 * i0: 2 * (i0 + icnt)]).to(dev)
    rounds = 7  # more rounds than slabs: the release flags sequence the reuse
    for rep in range(rounds):
        slab = rep % 2
        g.acquire(slab)
        out = g.slab(slab)[:cnt]
        out.fill_(float(rep))          # something else first, so that a stale slab cannot pass
        chain.run(seg, cnt, i0, out=out, n_in=icnt)
        g.submit(slab)
    g.finish()
    torch.cuda.synchronize()
    ok = 1
    if rank == 0:
        whole = chain.process_device(torch.from_numpy(x_host).to(dev)).cpu()
        for slab in range(2):
            got = torch.cat([g.result(slab, r)[: segs[r][1]].cpu() for r in range(world)])
            ok &= int(torch.equal(got.view(torch.int32), whole.view(torch.int32)))
        print("PEER_GATHER_OK" if ok else "PEER_GATHER_MISMATCH", g.stats(), flush=True)
    dist.barrier()
    g.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
