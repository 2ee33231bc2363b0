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


class _DevicePointer:
    """A raw device pointer presented through the CUDA array interface, so that torch can view library-owned memory."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class Gather:
    """ctypes front-end of b200sdr_gather (include/b200sdr/b200sdr.h): the gather of every rank's decimated audio to
    rank 0 over NCCL, slab by slab on a side stream.  The communicator, the slabs, the events and the grouped send/recv
    live in libb200sdr.so; the caller only distributes the 128-byte NCCL unique id (`unique_id()` on rank 0)."""

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from . import _native as N
        buf = C.create_string_buffer(128)
        N.check_status(N.lib.b200sdr_nccl_unique_id(buf), "b200sdr_nccl_unique_id")
        return buf.raw

    NCCL, PEER, PEER_COPY = 0, 1, 2

    def __init__(self, rank: int, world: int, floats_per_rank, slabs: int = 3, device: int = 0, unique_id: bytes | None = None,
                 mode: int = 0):
        import ctypes as C
        from . import _native as N
        self._N, self._C = N, C
        self.rank, self.world, self.slabs, self.device = rank, world, slabs, device
        self.floats_per_rank = [int(v) for v in floats_per_rank]
        assert len(self.floats_per_rank) == world
        self._counts = (C.c_size_t * world)(*self.floats_per_rank)
        self._id = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        cfg = N.GatherConfig()
        cfg.struct_size = C.sizeof(N.GatherConfig)
        cfg.rank, cfg.world, cfg.slabs, cfg.cuda_device = rank, world, slabs, device
        cfg.floats_per_rank = self._counts
        cfg.nccl_unique_id = C.cast(self._id, C.c_void_p) if self._id is not None else None
        cfg.mode = mode
        self.mode = mode if world > 1 else 0
        handle = C.c_void_p()
        N.check_status(N.lib.b200sdr_gather_create(C.byref(cfg), C.byref(handle)), "b200sdr_gather_create")
        self._h = handle

    def close(self):
        if getattr(self, "_h", None):
            self._N.lib.b200sdr_gather_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # peer mode: one exchange of IPC handles after construction (the caller all-gathers the blobs in rank order)
    def export_blob(self) -> bytes:
        size = self._N.lib.b200sdr_gather_exchange_size(self._h)
        buf = self._C.create_string_buffer(size)
        self._N.check_status(self._N.lib.b200sdr_gather_export(self._h, buf), "b200sdr_gather_export")
        return buf.raw

    def import_blobs(self, blobs: bytes) -> None:
        buf = self._C.create_string_buffer(blobs, len(blobs))
        self._N.check_status(self._N.lib.b200sdr_gather_import(self._h, buf), "b200sdr_gather_import")

    def slab(self, index: int):
        """This rank's part of slab `index` as a 1-D float32 torch tensor (a view of library-owned device memory)."""
        import torch
        ptr = self._N.lib.b200sdr_gather_slab(self._h, index)
        return torch.as_tensor(_DevicePointer(ptr, max(self.floats_per_rank[self.rank], 1)), device=torch.device("cuda", self.device))

    def result(self, index: int, rank: int):
        """Rank 0: rank `rank`'s part of gathered slab `index` (valid in stream order after finish())."""
        import torch
        ptr = self._N.lib.b200sdr_gather_result(self._h, index, rank)
        if not ptr:
            return None
        return torch.as_tensor(_DevicePointer(ptr, max(self.floats_per_rank[rank], 1)), device=torch.device("cuda", self.device))

    def _stream(self):
        import torch
        return torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream

    def acquire(self, index: int):
        self._N.check_status(self._N.lib.b200sdr_gather_acquire(self._h, index, self._stream()), "b200sdr_gather_acquire")

    def submit(self, index: int, floats_per_rank=None):
        counts = None
        if floats_per_rank is not None:
            counts = (self._C.c_size_t * self.world)(*[int(v) for v in floats_per_rank])
        self._N.check_status(self._N.lib.b200sdr_gather_submit(self._h, index, counts, self._stream()), "b200sdr_gather_submit")

    def finish(self):
        self._N.check_status(self._N.lib.b200sdr_gather_finish(self._h, self._stream()), "b200sdr_gather_finish")

    def stats(self):
        C = self._C
        gathers, moved, version = C.c_uint64(), C.c_uint64(), C.c_int32()
        self._N.lib.b200sdr_gather_stats(self._h, C.byref(gathers), C.byref(moved), C.byref(version))
        return {"gathers": gathers.value, "floats_moved": moved.value, "nccl_version": version.value}
