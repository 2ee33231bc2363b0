// b200sdr_gather -- the ONE exchange step of the multi-GPU decomposition: the decimated audio of
// every rank is gathered to rank 0 over NCCL (NVLink / NVSwitch), on a side stream, slab by slab, so that it overlaps the
// kernels of the next steps.  There is no collective on the filter path itself: every rank runs its own time segment
// (b200sdr_chain_segment / b200sdr_channelizer_segment) or its own channels.
//
// One object per rank (one process per GPU, or one host thread per GPU in a single process).  The caller distributes
// the 128-byte NCCL unique id from rank 0 to the other ranks by whatever transport it has (torch.distributed, MPI, a
// socket); everything else -- communicator, streams, events, slab ring, the grouped send/recv -- lives here.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): libb200sdr.so has no link-time dependency on it, single-GPU users
// never load it, and inside a process that already carries NCCL (PyTorch) the same library instance is used.
//
// Three transports:
//   mode 0 (NCCL)        one grouped ncclSend/ncclRecv per slab on the side stream.
//   mode 1 (peer store)  the gathered slabs of rank 0 are mapped into every rank (CUDA IPC over NVLink / NVSwitch peer memory) and
//                        b200sdr_gather_slab() hands the kernels THAT memory: the filter kernels store their audio straight into
//                        rank 0's buffer while they compute -- no kernel, no SM and no extra pass over the data.
//   mode 2 (peer copy)   the kernels write a LOCAL slab; the side stream moves it into rank 0's mapped slab with one
//                        cudaMemcpyAsync, i.e. a copy engine over NVLink -- no SM either, one extra read of the (decimated) audio.
//   Modes 1 and 2 sequence completion and buffer reuse with 32-bit flags written / awaited in stream order
//   (cuStreamWriteValue32 / cuStreamWaitValue32): rank r -> rank 0 "slab s holds round n of rank r", rank 0 -> rank r "slab s is
//   free again".
// Measured on 8 B200 (profiles/README.md, round 2): NCCL's send / receive kernels take SMs from the persistent filter kernel while a
// gather runs; peer stores leave the SMs alone but every kernel then ends with its remote stores still in flight (+5 % per C2
// step) and a bursty writer (the channelizer's audio kernel: 130 MB into rank 0 within 0.15 ms) is bound by rank 0's NVLink
// ingress; the copy engine spreads the same bytes over the whole next step.
#include <b200sdr/b200sdr.h>

#include <cuda.h>  // CUstream / CUdeviceptr / CUresult (types only; entry points are fetched at run time)
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

namespace b200sdr {
b200sdr_status chainFail(b200sdr_status status, const std::string& what);  // chain.cu: sets b200sdr_last_error()
}
using namespace b200sdr;

namespace {

// the handful of NCCL entry points used (stable since NCCL 2.7); types restated so that no NCCL header is needed
using ncclComm_t = struct ncclComm*;
struct NcclUniqueId {
  char internal[128];
};
constexpr int kNcclFloat = 7;  // ncclFloat32
struct Nccl {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool ok = false;
  std::string why;
};

const Nccl& nccl() {
  static Nccl api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* lib = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      api.why = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
    auto bind = [&](const char* sym) { return dlsym(lib, sym); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(bind("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(bind("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(bind("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(bind("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(bind("ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(bind("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(bind("ncclRecv"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(bind("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(bind("ncclGetVersion"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv;
    if (!api.ok) api.why = "libnccl.so.2 lacks a required entry point";
  });
  return api;
}

// stream-ordered 32-bit flag operations of the driver API, fetched at run time like the tensor-map encoder
using StreamValue32 = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
struct StreamOps {
  StreamValue32 write = nullptr, wait = nullptr;
};
const StreamOps& streamOps() {
  static StreamOps ops = [] {
    StreamOps o;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      o.write = reinterpret_cast<StreamValue32>(p);
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      o.wait = reinterpret_cast<StreamValue32>(p);
    (void)cudaGetLastError();
    return o;
  }();
  return ops;
}

b200sdr_status ncclFail(int rc, const char* where) {
  const Nccl& n = nccl();
  return chainFail(B200SDR_RUNTIME_ERROR, std::string(where) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error"));
}
b200sdr_status cudaFailG(cudaError_t e, const char* where) {
  return chainFail(e == cudaErrorMemoryAllocation ? B200SDR_OUT_OF_MEMORY : B200SDR_RUNTIME_ERROR, std::string(where) + ": " + cudaGetErrorString(e));
}

#define CUDA_OR_FAIL(call)                                 \
  do {                                                     \
    const cudaError_t e__ = (call);                        \
    if (e__ != cudaSuccess) return cudaFailG(e__, #call);  \
  } while (false)
#define NCCL_OR_FAIL(call)                          \
  do {                                              \
    const int rc__ = (call);                        \
    if (rc__ != 0) return ncclFail(rc__, #call);    \
  } while (false)

}  // namespace

struct b200sdr_gather {
  int device = 0, rank = 0, world = 1;
  unsigned slabs = 0;
  std::vector<size_t> floatsOf;  // [world] floats per rank per slab (capacity)
  std::vector<size_t> offsetOf;  // [world] float offset of each rank's part inside a gathered slab
  size_t gatheredFloats = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t side = nullptr;
  std::vector<float*> local;       // [slabs] this rank's part
  std::vector<float*> gathered;    // [slabs] rank 0 only: every rank's part
  std::vector<cudaEvent_t> filled, drained;
  std::vector<char> everSubmitted;
  uint64_t gathers = 0, floatsMoved = 0;
  // peer modes (1: kernels store into rank 0's slabs; 2: local slab + copy engine)
  int mode = 0;
  bool peerAny() const { return mode != 0; }
  bool imported = false;
  uint32_t* flags = nullptr;                 // this rank's flag block: [slabs][world] "filled" (used on rank 0) + [slabs] "free"
  std::vector<float*> remoteSlab;            // rank != 0: rank 0's gathered slabs mapped here
  std::vector<uint32_t*> remoteFlags;        // rank 0: every rank's flag block; rank != 0: [0] = rank 0's
  std::vector<uint32_t> round;               // [slabs] submissions so far
  uint32_t* filledFlag(unsigned slab, int r) const { return flags + slab * world + r; }
  uint32_t* freeFlagOf(uint32_t* block, unsigned slab) const { return block + slabs * world + slab; }
};

namespace {
constexpr size_t kIpcHandleBytes = 64;  // sizeof(cudaIpcMemHandle_t)
size_t exchangeBytes(const b200sdr_gather* g) { return (static_cast<size_t>(g->slabs) + 1u) * kIpcHandleBytes; }
}  // namespace

B200SDR_EXPORT b200sdr_status b200sdr_nccl_unique_id(void* id128) {
  if (!id128) return chainFail(B200SDR_INVALID_ARGUMENT, "id128 is null");
  const Nccl& n = nccl();
  if (!n.ok) return chainFail(B200SDR_RUNTIME_ERROR, n.why);
  NcclUniqueId id;
  NCCL_OR_FAIL(n.GetUniqueId(&id));
  std::memcpy(id128, id.internal, sizeof(id.internal));
  return B200SDR_OK;
}

B200SDR_EXPORT void b200sdr_gather_destroy(b200sdr_gather* g) {
  if (!g) return;
  DeviceGuard guard(g->device);
  if (g->side) cudaStreamSynchronize(g->side);
  if (g->comm) nccl().CommDestroy(g->comm);
  for (float* p : g->remoteSlab)
    if (p) cudaIpcCloseMemHandle(p);
  for (size_t r = 0; r < g->remoteFlags.size(); r++)
    if (g->remoteFlags[r] && g->remoteFlags[r] != g->flags) cudaIpcCloseMemHandle(g->remoteFlags[r]);
  if (g->flags) cudaFree(g->flags);
  for (float* p : g->local) cudaFree(p);
  for (float* p : g->gathered) cudaFree(p);
  for (cudaEvent_t e : g->filled) cudaEventDestroy(e);
  for (cudaEvent_t e : g->drained) cudaEventDestroy(e);
  if (g->side) cudaStreamDestroy(g->side);
  delete g;
}

B200SDR_EXPORT b200sdr_status b200sdr_gather_create(const b200sdr_gather_config* cfg, b200sdr_gather** out) {
  if (!cfg || !out) return chainFail(B200SDR_INVALID_ARGUMENT, "config and out must be non-null");
  *out = nullptr;
  if (cfg->struct_size != sizeof(b200sdr_gather_config)) return chainFail(B200SDR_INVALID_ARGUMENT, "struct_size mismatch");
  if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world || cfg->slabs < 2 || !cfg->floats_per_rank)
    return chainFail(B200SDR_INVALID_ARGUMENT, "need 0 <= rank < world, slabs >= 2 and floats_per_rank");
  if (cfg->mode > B200SDR_GATHER_PEER_COPY) return chainFail(B200SDR_INVALID_ARGUMENT, "unknown gather mode");
  const bool peer = cfg->mode != B200SDR_GATHER_NCCL && cfg->world > 1;
  if (cfg->world > 1 && !peer && !cfg->nccl_unique_id) return chainFail(B200SDR_INVALID_ARGUMENT, "nccl_unique_id is required when world > 1");
  if (peer && (!streamOps().write || !streamOps().wait)) return chainFail(B200SDR_RUNTIME_ERROR, "the driver lacks cuStreamWriteValue32 / cuStreamWaitValue32");
  b200sdr_gather* g = new (std::nothrow) b200sdr_gather();
  if (!g) return chainFail(B200SDR_OUT_OF_MEMORY, "host allocation failed");
  g->device = cfg->cuda_device;
  g->rank = cfg->rank;
  g->world = cfg->world;
  g->slabs = cfg->slabs;
  g->mode = peer ? static_cast<int>(cfg->mode) : 0;
  g->round.assign(cfg->slabs, 0u);
  g->floatsOf.assign(cfg->floats_per_rank, cfg->floats_per_rank + cfg->world);
  g->offsetOf.resize(cfg->world);
  for (int r = 0; r < cfg->world; r++) {
    g->offsetOf[r] = g->gatheredFloats;
    g->gatheredFloats += (g->floatsOf[r] + 63u) & ~static_cast<size_t>(63);  // every part starts on a 256-byte boundary
  }
  DeviceGuard guard(g->device);
  b200sdr_status st = B200SDR_OK;
  auto fail = [&](b200sdr_status s) {
    b200sdr_gather_destroy(g);
    return s;
  };
  if (guard.status != cudaSuccess) return fail(cudaFailG(guard.status, "cudaSetDevice"));
  cudaError_t e = cudaStreamCreateWithFlags(&g->side, cudaStreamNonBlocking);
  if (e != cudaSuccess) return fail(cudaFailG(e, "cudaStreamCreate"));
  for (unsigned s = 0; s < g->slabs && e == cudaSuccess; s++) {
    float* p = nullptr;
    e = cudaMalloc(reinterpret_cast<void**>(&p), sizeof(float) * (g->floatsOf[g->rank] ? g->floatsOf[g->rank] : 1));
    if (e == cudaSuccess) g->local.push_back(p);
    if (e == cudaSuccess && g->rank == 0) {
      e = cudaMalloc(reinterpret_cast<void**>(&p), sizeof(float) * (g->gatheredFloats ? g->gatheredFloats : 1));
      if (e == cudaSuccess) g->gathered.push_back(p);
    }
    cudaEvent_t ev = nullptr;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) g->filled.push_back(ev);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) g->drained.push_back(ev);
  }
  if (e != cudaSuccess) return fail(cudaFailG(e, "allocating the gather slabs"));
  g->everSubmitted.assign(g->slabs, 0);
  if (peer) {
    const size_t words = static_cast<size_t>(g->slabs) * g->world + g->slabs;
    e = cudaMalloc(reinterpret_cast<void**>(&g->flags), sizeof(uint32_t) * words);
    if (e == cudaSuccess) e = cudaMemset(g->flags, 0, sizeof(uint32_t) * words);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(cudaFailG(e, "allocating the gather flags"));
  }
  if (g->world > 1 && !peer) {
    const Nccl& n = nccl();
    if (!n.ok) return fail(chainFail(B200SDR_RUNTIME_ERROR, n.why));
    NcclUniqueId id;
    std::memcpy(id.internal, cfg->nccl_unique_id, sizeof(id.internal));
    const int rc = n.CommInitRank(&g->comm, g->world, id, g->rank);
    if (rc != 0) return fail(ncclFail(rc, "ncclCommInitRank"));
  }
  (void)st;
  *out = g;
  return B200SDR_OK;
}

B200SDR_EXPORT float* b200sdr_gather_slab(b200sdr_gather* g, uint32_t slab) {
  if (!g || slab >= g->slabs) return nullptr;
  if (g->mode == 0) return g->local[slab];
  if (!g->imported) return nullptr;  // the peer mappings exist after b200sdr_gather_import
  if (g->mode == 2 && g->rank != 0) return g->local[slab];  // moved by the copy engine at submit
  return (g->rank == 0 ? g->gathered[slab] : g->remoteSlab[slab]) + g->offsetOf[g->rank];  // rank 0's memory, this rank's part
}

B200SDR_EXPORT size_t b200sdr_gather_exchange_size(const b200sdr_gather* g) { return g && g->peerAny() ? exchangeBytes(g) : 0; }

// This rank's blob: the IPC handles of rank 0's gathered slabs (zero elsewhere) and of this rank's flag block.
B200SDR_EXPORT b200sdr_status b200sdr_gather_export(b200sdr_gather* g, void* blob) {
  if (!g || !blob || !g->peerAny()) return chainFail(B200SDR_INVALID_ARGUMENT, "not a peer-mode gather");
  DeviceGuard guard(g->device);
  unsigned char* out = static_cast<unsigned char*>(blob);
  std::memset(out, 0, exchangeBytes(g));
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == kIpcHandleBytes, "cudaIpcMemHandle_t is 64 bytes");
  if (g->rank == 0)
    for (unsigned s = 0; s < g->slabs; s++) {
      CUDA_OR_FAIL(cudaIpcGetMemHandle(&h, g->gathered[s]));
      std::memcpy(out + s * kIpcHandleBytes, &h, sizeof(h));
    }
  CUDA_OR_FAIL(cudaIpcGetMemHandle(&h, g->flags));
  std::memcpy(out + g->slabs * kIpcHandleBytes, &h, sizeof(h));
  return B200SDR_OK;
}

// blobs: `world` blobs of b200sdr_gather_exchange_size() bytes each, in rank order (gathered by the caller's transport)
B200SDR_EXPORT b200sdr_status b200sdr_gather_import(b200sdr_gather* g, const void* blobs) {
  if (!g || !blobs || !g->peerAny()) return chainFail(B200SDR_INVALID_ARGUMENT, "not a peer-mode gather");
  if (g->imported) return B200SDR_OK;
  DeviceGuard guard(g->device);
  const unsigned char* in = static_cast<const unsigned char*>(blobs);
  const size_t each = exchangeBytes(g);
  cudaIpcMemHandle_t h;
  g->remoteFlags.assign(g->world, nullptr);
  g->remoteFlags[g->rank] = g->flags;
  if (g->rank == 0) {
    for (int r = 1; r < g->world; r++) {  // every rank's flag block: rank 0 writes their "free" flags
      std::memcpy(&h, in + r * each + g->slabs * kIpcHandleBytes, sizeof(h));
      void* p = nullptr;
      CUDA_OR_FAIL(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      g->remoteFlags[r] = static_cast<uint32_t*>(p);
    }
  } else {
    g->remoteSlab.assign(g->slabs, nullptr);
    for (unsigned s = 0; s < g->slabs; s++) {  // rank 0's gathered slabs: the kernels of this rank store into them
      std::memcpy(&h, in + s * kIpcHandleBytes, sizeof(h));
      void* p = nullptr;
      CUDA_OR_FAIL(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      g->remoteSlab[s] = static_cast<float*>(p);
    }
    std::memcpy(&h, in + g->slabs * kIpcHandleBytes, sizeof(h));
    void* p = nullptr;
    CUDA_OR_FAIL(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g->remoteFlags[0] = static_cast<uint32_t*>(p);
  }
  g->imported = true;
  return B200SDR_OK;
}

B200SDR_EXPORT b200sdr_status b200sdr_gather_acquire(b200sdr_gather* g, uint32_t slab, cudaStream_t stream) {
  if (!g || slab >= g->slabs) return chainFail(B200SDR_INVALID_ARGUMENT, "bad slab");
  if (!g->everSubmitted[slab]) return B200SDR_OK;
  DeviceGuard guard(g->device);
  if (g->mode == 1 && g->rank != 0) {
    // rank 0 has released round `round[slab]` of this slab (a flag in THIS rank's memory, written by rank 0 in stream order)
    if (streamOps().wait(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(g->freeFlagOf(g->flags, slab)), g->round[slab],
                         CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
      return chainFail(B200SDR_RUNTIME_ERROR, "cuStreamWaitValue32 failed");
    return B200SDR_OK;
  }
  CUDA_OR_FAIL(cudaStreamWaitEvent(stream, g->drained[slab], 0));  // the slab's previous gather has read it
  return B200SDR_OK;
}

B200SDR_EXPORT b200sdr_status b200sdr_gather_submit(b200sdr_gather* g, uint32_t slab, const size_t* floatsPerRank, cudaStream_t stream) {
  if (!g || slab >= g->slabs) return chainFail(B200SDR_INVALID_ARGUMENT, "bad slab");
  DeviceGuard guard(g->device);
  const size_t* counts = floatsPerRank ? floatsPerRank : g->floatsOf.data();
  for (int r = 0; r < g->world; r++)
    if (counts[r] > g->floatsOf[r]) return chainFail(B200SDR_OUT_OF_RANGE, "more floats than the slab holds");
  if (g->mode == 2 && g->rank != 0) {
    if (!g->imported) return chainFail(B200SDR_INVALID_STATE, "b200sdr_gather_import has not been called");
    const uint32_t round = ++g->round[slab];
    CUDA_OR_FAIL(cudaEventRecord(g->filled[slab], stream));
    CUDA_OR_FAIL(cudaStreamWaitEvent(g->side, g->filled[slab], 0));
    // rank 0 has released the previous round of its slab (a flag in THIS rank's memory); back-pressure stays on the side stream
    if (round > 1 && streamOps().wait(reinterpret_cast<CUstream>(g->side), reinterpret_cast<CUdeviceptr>(g->freeFlagOf(g->flags, slab)), round - 1,
                                      CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
      return chainFail(B200SDR_RUNTIME_ERROR, "cuStreamWaitValue32 failed");
    if (counts[g->rank])  // a copy engine moves the slab over NVLink into rank 0's memory
      CUDA_OR_FAIL(cudaMemcpyAsync(g->remoteSlab[slab] + g->offsetOf[g->rank], g->local[slab], sizeof(float) * counts[g->rank], cudaMemcpyDefault, g->side));
    if (streamOps().write(reinterpret_cast<CUstream>(g->side), reinterpret_cast<CUdeviceptr>(g->remoteFlags[0] + slab * g->world + g->rank), round, 0) !=
        CUDA_SUCCESS)
      return chainFail(B200SDR_RUNTIME_ERROR, "cuStreamWriteValue32 failed");
    CUDA_OR_FAIL(cudaEventRecord(g->drained[slab], g->side));  // the local slab may be written again
    g->everSubmitted[slab] = 1;
    g->gathers++;
    g->floatsMoved += counts[g->rank];
    return B200SDR_OK;
  }
  if (g->peerAny()) {
    if (!g->imported) return chainFail(B200SDR_INVALID_STATE, "b200sdr_gather_import has not been called");
    const uint32_t round = ++g->round[slab];
    // "slab holds round `round` of this rank": a flag in RANK 0's memory, written behind the kernels that stored the audio there
    uint32_t* mine = g->rank == 0 ? g->filledFlag(slab, 0) : g->remoteFlags[0] + slab * g->world + g->rank;
    if (streamOps().write(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(mine), round, 0) != CUDA_SUCCESS)
      return chainFail(B200SDR_RUNTIME_ERROR, "cuStreamWriteValue32 failed");
    if (g->rank == 0) {
      // the side stream collects every rank's flag, marks the slab complete and hands it back to the senders
      for (int r = 0; r < g->world; r++)
        if (streamOps().wait(reinterpret_cast<CUstream>(g->side), reinterpret_cast<CUdeviceptr>(g->filledFlag(slab, r)), round, CU_STREAM_WAIT_VALUE_GEQ) !=
            CUDA_SUCCESS)
          return chainFail(B200SDR_RUNTIME_ERROR, "cuStreamWaitValue32 failed");
      CUDA_OR_FAIL(cudaEventRecord(g->drained[slab], g->side));
      for (int r = 1; r < g->world; r++)
        if (streamOps().write(reinterpret_cast<CUstream>(g->side), reinterpret_cast<CUdeviceptr>(g->freeFlagOf(g->remoteFlags[r], slab)), round, 0) !=
            CUDA_SUCCESS)
          return chainFail(B200SDR_RUNTIME_ERROR, "cuStreamWriteValue32 failed");
    }
    g->everSubmitted[slab] = 1;
    g->gathers++;
    for (int r = 0; r < g->world; r++) g->floatsMoved += (g->rank == 0 || r == g->rank) ? counts[r] : 0;
    return B200SDR_OK;
  }
  CUDA_OR_FAIL(cudaEventRecord(g->filled[slab], stream));
  CUDA_OR_FAIL(cudaStreamWaitEvent(g->side, g->filled[slab], 0));
  if (g->rank == 0 && counts[0])
    CUDA_OR_FAIL(cudaMemcpyAsync(g->gathered[slab] + g->offsetOf[0], g->local[slab], sizeof(float) * counts[0], cudaMemcpyDeviceToDevice, g->side));
  if (g->world > 1) {
    const Nccl& n = nccl();
    NCCL_OR_FAIL(n.GroupStart());
    if (g->rank == 0) {
      for (int r = 1; r < g->world; r++)
        if (counts[r]) NCCL_OR_FAIL(n.Recv(g->gathered[slab] + g->offsetOf[r], counts[r], kNcclFloat, r, g->comm, g->side));
    } else if (counts[g->rank]) {
      NCCL_OR_FAIL(n.Send(g->local[slab], counts[g->rank], kNcclFloat, 0, g->comm, g->side));
    }
    NCCL_OR_FAIL(n.GroupEnd());
  }
  CUDA_OR_FAIL(cudaEventRecord(g->drained[slab], g->side));
  g->everSubmitted[slab] = 1;
  g->gathers++;
  for (int r = 0; r < g->world; r++) g->floatsMoved += (g->rank == 0 || r == g->rank) ? counts[r] : 0;
  return B200SDR_OK;
}

B200SDR_EXPORT b200sdr_status b200sdr_gather_finish(b200sdr_gather* g, cudaStream_t stream) {
  if (!g) return chainFail(B200SDR_INVALID_ARGUMENT, "gather is null");
  DeviceGuard guard(g->device);
  for (unsigned s = 0; s < g->slabs; s++) {
    if (!g->everSubmitted[s]) continue;
    if (g->mode == 1 && g->rank != 0) {  // delivered and released by rank 0
      if (streamOps().wait(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(g->freeFlagOf(g->flags, s)), g->round[s],
                           CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
        return chainFail(B200SDR_RUNTIME_ERROR, "cuStreamWaitValue32 failed");
    } else {
      CUDA_OR_FAIL(cudaStreamWaitEvent(stream, g->drained[s], 0));
    }
  }
  return B200SDR_OK;
}

B200SDR_EXPORT const float* b200sdr_gather_result(const b200sdr_gather* g, uint32_t slab, int32_t rank) {
  if (!g || g->rank != 0 || slab >= g->slabs || rank < 0 || rank >= g->world) return nullptr;
  return g->gathered[slab] + g->offsetOf[rank];
}

B200SDR_EXPORT void b200sdr_gather_stats(const b200sdr_gather* g, uint64_t* gathers, uint64_t* floatsMoved, int32_t* ncclVersion) {
  if (gathers) *gathers = g ? g->gathers : 0;
  if (floatsMoved) *floatsMoved = g ? g->floatsMoved : 0;
  if (ncclVersion) {
    *ncclVersion = 0;
    if (g && g->world > 1 && nccl().GetVersion) {
      int v = 0;
      if (nccl().GetVersion(&v) == 0) *ncclVersion = v;
    }
  }
}
