"""Multi-GPU partitioning of the hot path: no collective on the filter path.

  * by channel  -- channel c runs on rank c mod G (the wideband input is replicated);
  * by time     -- rank g computes final outputs [M_g, M_g+1) of one channel and reads the input
                   segment those outputs need (a look-ahead halo of window - stride samples).

The only exchange is the gather of the decimated outputs to rank 0 (NCCL on GPUs, gloo in CPU tests).
The segment arithmetic here is pure Python so that it can be tested without a GPU; it mirrors
b200sdr_chain_segment() in csrc/chain.cu (tests check both agree)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Segment:
    first_output: int
    output_count: int
    first_input: int
    input_count: int


def chain_stride(rf_decim: int, audio_decim: int, has_audio_fir: bool) -> int:
    return max(1, rf_decim) * (max(1, audio_decim) if has_audio_fir else 1)


def chain_window(rf_taps: int, rf_decim: int, audio_taps: int, fm: bool, has_audio_fir: bool) -> int:
    """Input samples one final output needs: ((T2-1) + fm) * D1 + T1 (S)."""
    return ((audio_taps - 1 if has_audio_fir else 0) + (1 if fm else 0)) * max(1, rf_decim) + rf_taps


def time_segment(num_outputs: int, parts: int, index: int, stride: int, window: int) -> Segment:
    if parts <= 0 or not (0 <= index < parts):
        raise ValueError("bad segment request")
    base, extra = divmod(num_outputs, parts)
    first = index * base + min(index, extra)
    count = base + (1 if index < extra else 0)
    return Segment(first, count, first * stride, 0 if count == 0 else (count - 1) * stride + window)


def channels_of_rank(num_channels: int, world_size: int, rank: int) -> list[int]:
    """Interleaved assignment so AM/FM cost is balanced."""
    return list(range(rank, num_channels, world_size))


def gather_to_rank0(local, counts: list[int], group=None):
    """Gather variable-length 1-D tensors to rank 0 with torch.distributed (send/recv, no padding).
    Returns the concatenated tensor on rank 0 and None elsewhere."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    assert len(counts) == world and local.numel() == counts[rank]
    if world == 1:
        return local
    if rank == 0:
        out = torch.empty(sum(counts), dtype=local.dtype, device=local.device)
        out[: counts[0]] = local
        offset, reqs = counts[0], []
        for src in range(1, world):
            if counts[src]:
                reqs.append(dist.irecv(out[offset: offset + counts[src]], src=src, group=group))
            offset += counts[src]
        for r in reqs:
            r.wait()
        return out
    if counts[rank]:
        dist.send(local.contiguous(), dst=0, group=group)
    return None
