// The single-op graph nodes of the hot path behind the Filter / Source / Sink contract (abi/nodes.h), each enqueuing
// ONE call into the sm_100a kernels of libb200sdr.so (include/gsdr/gsdr.h) on its queue's stream:
//
//   Int8ToFloat, CosineSource (real/complex), MultiplyCcc, Fir (FF/FC/CC/CF, decimating), QuadAmDemod, QuadFmDemod,
//   Magnitude, AddConst, AddConstToVectorLength, CudaMemcpyFilter, FileReader, port remapping, byte-count monitor.
//
// Sample-count, consumption and state rules follow the reference's wrappers (cited at each node).  What is different
// is the stream-state carrier: PortInput keeps a port's unconsumed input in place -- consuming is pointer arithmetic,
// the remainder is moved to the front of the allocation only when a request no longer fits behind it -- where the
// reference copies the remainder into a twin allocation after EVERY readOutput (BaseSink.cpp:150-170 ->
// RelocatableResizableBuffer.cpp:79-103).
#include <gsdr/conversion.h>
#include <gsdr/gsdr.h>

#include <cmath>
#include <cstring>
#include <map>
#include <memory>

#include "internal.h"
#include "json_min.h"
#include "port_input.h"

namespace gs {

// ---------------------------------------------------------------------------------------------------------------
// PortInput
// ---------------------------------------------------------------------------------------------------------------
namespace {

class ViewBuffer final : public IBuffer {
 public:
  ViewBuffer(IMemory* keepAlive, uint8_t* base, IBufferRangeMutableCapacity* range) noexcept : mMemory(keepAlive), mBase(base), mRange(range) {}
  uint8_t* base() noexcept final { return mBase; }
  const uint8_t* base() const noexcept final { return mBase; }
  IBufferRange* range() noexcept final { return mRange.get(); }
  const IBufferRange* range() const noexcept final { return mRange.get(); }

 private:
  ConstRef<IMemory> mMemory;
  uint8_t* const mBase;
  ConstRef<IBufferRangeMutableCapacity> mRange;
  REF_COUNTED(ViewBuffer);
};

}  // namespace

PortInput::PortInput(IAllocator* allocator, const IBufferCopier* copier, ICudaCommandQueue* queue, bool host) noexcept
    : mAllocator(allocator), mCopier(copier), mQueue(queue), mHost(host) {}

PortInput::~PortInput() noexcept {
  if (mFence || mOtherFence) {
    CudaDevicePushPop device(mQueue->cudaDevice());
    if (mFence) cudaEventDestroy(mFence);
    if (mOtherFence) cudaEventDestroy(mOtherFence);
  }
}

Status PortInput::makeRoom(size_t bytes) noexcept {
  const size_t used = mEnd - mOffset;
  if (mMemory != nullptr && mCapacity - mEnd >= bytes) return Status_Success;
  if (mMemory != nullptr && used + bytes <= mCapacity && mOffset >= used) {
    // compact in place: source and destination do not overlap.  On a pinned host port the bytes in front of mOffset may
    // still be read by a queued host->device copy (a partial drain leaves the remainder here): wait for that copy before
    // the host overwrites them
    if (mHost && mFencePending) {
      CUDA_DEV_PUSH_POP_OR_RET_STATUS(mQueue->cudaDevice());
      SAFE_CUDA_OR_RET_STATUS(cudaEventSynchronize(mFence));
      mFencePending = false;
    }
    FWD_IF_ERR(mCopier->copy(mMemory.get()->data(), mMemory.get()->data() + mOffset, used));
    mOffset = 0;
    mEnd = used;
    return Status_Success;
  }
  size_t wanted = used + bytes;
  if (wanted < 2 * mCapacity) wanted = 2 * mCapacity;  // amortised growth; also leaves room to compact without overlap
  if (wanted < 8192) wanted = 8192;                    // reference default input buffer size (BaseSink.cpp:50)
  Ref<IMemory> bigger;
  UNWRAP_OR_FWD_STATUS(bigger, mAllocator->allocate(roundUp(wanted, 256)));
  if (used) FWD_IF_ERR(mCopier->copy(bigger.get()->data(), mMemory.get()->data() + mOffset, used));
  if (mHost && mFencePending) {  // the old pinned block may still be read by a queued copy
    CUDA_DEV_PUSH_POP_OR_RET_STATUS(mQueue->cudaDevice());
    SAFE_CUDA_OR_RET_STATUS(cudaEventSynchronize(mFence));
    mFencePending = false;
  }
  mMemory = bigger;
  mCapacity = bigger.get()->capacity();
  mOffset = 0;
  mEnd = used;
  return Status_Success;
}

Result<IBuffer> PortInput::request(size_t bytes) noexcept {
  GS_REQUIRE_OR_RET_RESULT(!mCheckedOut, "Cannot request buffer - it is already checked out");
  if (mHost && mEnd == mOffset && mFencePending) {
    // everything was drained by a copy that may still be running: continue in the other pinned block
    std::swap(mFence, mOtherFence);
    std::swap(mFencePending, mOtherFencePending);
    Ref<IMemory> tmp = mMemory;
    mMemory = mOther;
    mOther = tmp;
    mCapacity = mMemory != nullptr ? mMemory.get()->capacity() : 0;
    mOffset = mEnd = 0;
    if (mFencePending) {
      CUDA_DEV_PUSH_POP_OR_RET_RESULT(mQueue->cudaDevice());
      SAFE_CUDA_OR_RET_RESULT(cudaEventSynchronize(mFence));
      mFencePending = false;
    }
  }
  FWD_IN_RESULT_IF_ERR(makeRoom(bytes));
  Ref<IBufferRangeMutableCapacity> range;
  UNWRAP_OR_FWD_RESULT(range, newBufferRange());
  range.get()->setCapacity(mCapacity - mEnd);
  IBuffer* view = new (std::nothrow) ViewBuffer(mMemory.get(), mMemory.get()->data() + mEnd, range.get());
  if (view != nullptr) mCheckedOut = true;
  return makeRefResultNonNull(view);
}

Status PortInput::commit(size_t bytes) noexcept {
  GS_REQUIRE_OR_RET_STATUS(mCheckedOut, "Cannot commit buffer - it was not checked out");
  GS_REQUIRE_OR_RET_STATUS(bytes <= mCapacity - mEnd, "Cannot commit buffer - the committed number of bytes exceeds its capacity");
  mEnd += bytes;
  mCheckedOut = false;
  return Status_Success;
}

void PortInput::consume(size_t bytes) noexcept {
  mOffset += bytes < mEnd - mOffset ? bytes : mEnd - mOffset;
  if (mOffset == mEnd) mOffset = mEnd = 0;
}

Status PortInput::fenceDrained() noexcept {
  if (!mHost) return Status_Success;
  CUDA_DEV_PUSH_POP_OR_RET_STATUS(mQueue->cudaDevice());
  if (mFence == nullptr) SAFE_CUDA_OR_RET_STATUS(cudaEventCreateWithFlags(&mFence, cudaEventDisableTiming));
  SAFE_CUDA_OR_RET_STATUS(cudaEventRecord(mFence, mQueue->cudaStream()));
  mFencePending = true;
  return Status_Success;
}

namespace {

// ---------------------------------------------------------------------------------------------------------------
// Common part of every GPU node: queue, input ports, output copier.
// ---------------------------------------------------------------------------------------------------------------
struct NodeParts {
  ConstRef<ICudaCommandQueue> queue;
  Ref<IAllocator> allocator;
  Ref<IBufferCopier> d2d;
  std::vector<std::unique_ptr<PortInput>> ports;

  explicit NodeParts(ICudaCommandQueue* q) : queue(q) {}
  Status init(IFactories* f, size_t portCount, bool hostInput = false) noexcept {
    UNWRAP_OR_FWD_STATUS(allocator, f->getCudaAllocatorFactory()->createCudaAllocator(queue, 256, hostInput));
    UNWRAP_OR_FWD_STATUS(d2d, f->getCudaBufferCopierFactory()->createBufferCopier(queue, cudaMemcpyDeviceToDevice));
    Ref<IBufferCopier> portCopier = d2d;
    if (hostInput) portCopier = f->getSysMemCopier();  // pinned host memory is compacted / grown with memmove
    try {
      for (size_t i = 0; i < portCount; i++) ports.emplace_back(new PortInput(allocator.get(), portCopier.get(), queue, hostInput));
    }
    IF_CATCH_RETURN_STATUS
    return Status_Success;
  }
  PortInput& port(size_t i) const noexcept { return *ports[i]; }
  int32_t device() const noexcept { return queue->cudaDevice(); }
  cudaStream_t stream() const noexcept { return queue->cudaStream(); }
};

#define GS_SINK_METHODS(parts__, preferred__)                                                                    \
  Result<IBuffer> requestBuffer(size_t port, size_t byteCount) noexcept final {                                    \
    GS_REQUIRE_OR_RET_RESULT_FMT(port < parts__.ports.size(), "Cannot request buffer. Input port [%zu] is out of range.", port); \
    return parts__.port(port).request(byteCount);                                                                \
  }                                                                                                               \
  Status commitBuffer(size_t port, size_t byteCount) noexcept final {                                              \
    GS_REQUIRE_OR_RET_STATUS_FMT(port < parts__.ports.size(), "Cannot commit buffer. Input port [%zu] is out of range", port); \
    return parts__.port(port).commit(byteCount);                                                                 \
  }                                                                                                               \
  size_t preferredInputBufferSize(size_t) noexcept final { return preferred__; }

#define GS_SOURCE_COMMON(parts__, alignment__)                                        \
  size_t getOutputSizeAlignment(size_t port) noexcept final { return port == 0 ? (alignment__) : 0; } \
  IBufferCopier* getOutputCopier(size_t port) noexcept final { return port == 0 ? parts__.d2d.get().get() : nullptr; }

// The reference's nodes check `portCount != 0`; its Int8ToFloat checks `0 == portCount` (Int8ToFloat.cpp:81), so callers
// of the reference must pass 0 there.  Every node of this library accepts both spellings as long as bufs[0] is given.
#define GS_REQUIRE_OUTPUT(bufs__) GS_REQUIRE_OR_RET_STATUS((bufs__) != nullptr && (bufs__)[0] != nullptr, "One output port is required")

constexpr size_t kStep = size_t(1) << 20;  // preferredInputBufferSize of the reference's nodes (Fir.cpp:311 etc.)

// ---------------------------------------------------------------------------------------------------------------
// Element-wise nodes: n = min(available, room), one launch, consume what was used.
// ---------------------------------------------------------------------------------------------------------------
enum class MapKind { Int8ToFloat, QuadAm, Magnitude, QuadFm, AddConst, AddToMagnitude };

class MapFilter final : public Filter {
 public:
  static Result<Filter> create(MapKind kind, float param, ICudaCommandQueue* queue, IFactories* f) noexcept {
    NON_NULL_PARAM_OR_RET(queue);
    MapFilter* node = new (std::nothrow) MapFilter(kind, param, queue);
    NON_NULL_OR_RET(node);
    const Status st = node->mParts.init(f, 1);
    if (st != Status_Success) {
      node->unref();  // floating: never reffed, so this destroys it
      return ERR_RESULT(st);
    }
    return makeRefResultNonNull<Filter>(node);
  }

 private:
  MapFilter(MapKind kind, float param, ICudaCommandQueue* queue) noexcept : mKind(kind), mParam(param), mParts(queue) {
    switch (kind) {
      case MapKind::Int8ToFloat: mInBytes = 1; mOutBytes = 4; break;                       // scalars: I and Q separately
      case MapKind::QuadAm: case MapKind::Magnitude: case MapKind::QuadFm: mInBytes = 8; mOutBytes = 4; break;
      case MapKind::AddConst: mInBytes = 4; mOutBytes = 4; break;
      case MapKind::AddToMagnitude: mInBytes = 8; mOutBytes = 8; break;
    }
  }
  size_t available() const noexcept {
    const size_t n = mParts.port(0).used() / mInBytes;
    return mKind == MapKind::QuadFm ? (n == 0 ? 0 : n - 1) : n;  // the discriminator keeps one sample (QuadFmDemod.cpp:76-84)
  }

 public:
  GS_SINK_METHODS(mParts, kStep)
  GS_SOURCE_COMMON(mParts, 32 * mOutBytes)
  size_t getOutputDataSize(size_t port) noexcept final { return port == 0 ? available() * mOutBytes : 0; }

  Status readOutput(IBuffer** bufs, size_t) noexcept final {
    GS_REQUIRE_OUTPUT(bufs);
    IBuffer* out = bufs[0];
    size_t n = available();
    const size_t room = out->range()->remaining() / mOutBytes;
    if (n > room) n = room;
    if (n == 0) return Status_Success;
    const void* in = mParts.port(0).data();
    void* dst = out->writePtr();
    cudaError_t e = cudaSuccess;
    switch (mKind) {
      case MapKind::Int8ToFloat:
        e = gsdrInt8ToNormFloat(static_cast<const int8_t*>(in), static_cast<float*>(dst), n, mParts.device(), mParts.stream());
        break;
      case MapKind::QuadAm:
        e = gsdrQuadAmDemod(static_cast<const cuComplex*>(in), static_cast<float*>(dst), n, mParts.device(), mParts.stream());
        break;
      case MapKind::Magnitude:
        e = gsdrMagnitude(static_cast<const cuComplex*>(in), static_cast<float*>(dst), n, mParts.device(), mParts.stream());
        break;
      case MapKind::QuadFm:
        e = gsdrQuadFmDemod(static_cast<const cuComplex*>(in), static_cast<float*>(dst), mParam, n, mParts.device(), mParts.stream());
        break;
      case MapKind::AddConst:
        e = gsdrAddConstFF(static_cast<const float*>(in), mParam, static_cast<float*>(dst), n, mParts.device(), mParts.stream());
        break;
      case MapKind::AddToMagnitude:
        e = gsdrAddToMagnitude(static_cast<const cuComplex*>(in), mParam, static_cast<cuComplex*>(dst), n, mParts.device(), mParts.stream());
        break;
    }
    SAFE_CUDA_OR_RET_STATUS(e);
    FWD_IF_ERR(out->range()->increaseEndOffset(n * mOutBytes));
    mParts.port(0).consume(n * mInBytes);
    return Status_Success;
  }

 private:
  const MapKind mKind;
  const float mParam;
  size_t mInBytes = 1, mOutBytes = 1;
  NodeParts mParts;
  REF_COUNTED(MapFilter);
};

// ---------------------------------------------------------------------------------------------------------------
// MultiplyCcc: out[i] = port0[i] * port1[i]   (reference Multiply.cpp:70-159)
// ---------------------------------------------------------------------------------------------------------------
class MultiplyFilter final : public Filter {
 public:
  static Result<Filter> create(ICudaCommandQueue* queue, IFactories* f) noexcept {
    NON_NULL_PARAM_OR_RET(queue);
    MultiplyFilter* node = new (std::nothrow) MultiplyFilter(queue);
    NON_NULL_OR_RET(node);
    const Status st = node->mParts.init(f, 2);
    if (st != Status_Success) {
      node->unref();
      return ERR_RESULT(st);
    }
    return makeRefResultNonNull<Filter>(node);
  }

  Result<IBuffer> requestBuffer(size_t port, size_t byteCount) noexcept final {
    GS_REQUIRE_OR_RET_RESULT_FMT(port < 2, "Cannot request buffer. Input port [%zu] is out of range.", port);
    return mParts.port(port).request(byteCount);
  }
  Status commitBuffer(size_t port, size_t byteCount) noexcept final {
    GS_REQUIRE_OR_RET_STATUS_FMT(port < 2, "Cannot commit buffer. Input port [%zu] is out of range", port);
    return mParts.port(port).commit(byteCount);
  }
  // ask for exactly what the lagging port is missing (Multiply.cpp:92-130)
  size_t preferredInputBufferSize(size_t port) noexcept final {
    const size_t u0 = mParts.port(0).used(), u1 = mParts.port(1).used();
    constexpr size_t maxSize = size_t(100) << 20;
    if (port > 1) return 0;
    if (u0 == 0 && u1 == 0) return 8192 * sizeof(cuComplex);
    const size_t mine = port == 0 ? u0 : u1, other = port == 0 ? u1 : u0;
    return mine >= other ? 0 : (other - mine < maxSize ? other - mine : maxSize);
  }
  GS_SOURCE_COMMON(mParts, 32 * sizeof(cuComplex))
  size_t getOutputDataSize(size_t port) noexcept final { return port == 0 ? available() * sizeof(cuComplex) : 0; }

  Status readOutput(IBuffer** bufs, size_t) noexcept final {
    GS_REQUIRE_OUTPUT(bufs);
    IBuffer* out = bufs[0];
    size_t n = available();
    const size_t room = out->range()->remaining() / sizeof(cuComplex);
    if (n > room) n = room;
    if (n == 0) return Status_Success;
    SAFE_CUDA_OR_RET_STATUS(gsdrMultiplyCC(reinterpret_cast<const cuComplex*>(mParts.port(0).data()),
                                           reinterpret_cast<const cuComplex*>(mParts.port(1).data()), out->writePtr<cuComplex>(), n,
                                           mParts.device(), mParts.stream()));
    FWD_IF_ERR(out->range()->increaseEndOffset(n * sizeof(cuComplex)));
    mParts.port(0).consume(n * sizeof(cuComplex));
    mParts.port(1).consume(n * sizeof(cuComplex));
    return Status_Success;
  }

 private:
  explicit MultiplyFilter(ICudaCommandQueue* queue) noexcept : mParts(queue) {}
  size_t available() const noexcept {
    const size_t a = mParts.port(0).used() / sizeof(cuComplex), b = mParts.port(1).used() / sizeof(cuComplex);
    return a < b ? a : b;
  }
  NodeParts mParts;
  REF_COUNTED(MultiplyFilter);
};

// ---------------------------------------------------------------------------------------------------------------
// Fir: decimating FIR, taps used as given (reference Fir.cpp:47-278)
// ---------------------------------------------------------------------------------------------------------------
class FirFilter final : public Filter {
 public:
  static Result<Filter> create(SampleType tapType, SampleType elementType, size_t decimation, const float* taps, size_t tapCount,
                               ICudaCommandQueue* queue, IFactories* f) noexcept {
    NON_NULL_PARAM_OR_RET(queue);
    NON_NULL_PARAM_OR_RET(taps);
    GS_REQUIRE_OR_RET_RESULT(tapCount > 0, "A FIR needs at least one tap");
    GS_REQUIRE_OR_RET_RESULT_FMT(tapType == SampleType_Float || tapType == SampleType_FloatComplex, "Unsupported tap type [%u]", tapType);
    GS_REQUIRE_OR_RET_RESULT_FMT(elementType == SampleType_Float || elementType == SampleType_FloatComplex, "Unsupported element type [%u]",
                                 elementType);
    FirFilter* node = new (std::nothrow) FirFilter(tapType, elementType, decimation == 0 ? 1 : decimation, tapCount, queue);  // Fir.cpp:119
    NON_NULL_OR_RET(node);
    Status st = node->mParts.init(f, 1);
    if (st == Status_Success) st = node->uploadTaps(taps);
    if (st != Status_Success) {
      node->unref();
      return ERR_RESULT(st);
    }
    return makeRefResultNonNull<Filter>(node);
  }

  GS_SINK_METHODS(mParts, kStep)
  GS_SOURCE_COMMON(mParts, 32 * mOutBytes)
  size_t getOutputDataSize(size_t port) noexcept final { return port == 0 ? numOutputs() * mOutBytes : 0; }

  Status readOutput(IBuffer** bufs, size_t) noexcept final {
    GS_REQUIRE_OUTPUT(bufs);
    IBuffer* out = bufs[0];
    size_t n = numOutputs();
    const size_t room = out->range()->remaining() / mOutBytes;
    if (n > room) n = room;
    if (n == 0) return Status_Success;
    const void* in = mParts.port(0).data();
    const float* taps = mTaps.get()->as<float>();
    cudaError_t e;
    if (mTapType == SampleType_Float && mElemType == SampleType_Float) {
      e = gsdrFirFF(mD, taps, mT, static_cast<const float*>(in), out->writePtr<float>(), n, mParts.device(), mParts.stream());
    } else if (mTapType == SampleType_Float) {
      e = gsdrFirFC(mD, taps, mT, static_cast<const cuComplex*>(in), out->writePtr<cuComplex>(), n, mParts.device(), mParts.stream());
    } else if (mElemType == SampleType_FloatComplex) {
      e = gsdrFirCC(mD, reinterpret_cast<const cuComplex*>(taps), mT, static_cast<const cuComplex*>(in), out->writePtr<cuComplex>(), n,
                    mParts.device(), mParts.stream());
    } else {
      e = gsdrFirCF(mD, reinterpret_cast<const cuComplex*>(taps), mT, static_cast<const float*>(in), out->writePtr<cuComplex>(), n,
                    mParts.device(), mParts.stream());
    }
    SAFE_CUDA_OR_RET_STATUS(e);
    FWD_IF_ERR(out->range()->increaseEndOffset(n * mOutBytes));
    mParts.port(0).consume(n * mD * mInBytes);  // exactly nOut*D elements: the T-1 history stays (Fir.cpp:274-276)
    return Status_Success;
  }

 private:
  FirFilter(SampleType tapType, SampleType elemType, size_t D, size_t T, ICudaCommandQueue* queue) noexcept
      : mTapType(tapType), mElemType(elemType), mD(D), mT(T), mParts(queue) {
    mInBytes = elemType == SampleType_Float ? 4 : 8;
    mOutBytes = (elemType == SampleType_Float && tapType == SampleType_Float) ? 4 : 8;
  }
  // `taps` holds tapCount floats, or tapCount (re, im) pairs for complex taps.  (The reference allocates room for
  // tapCount floats in both cases, Fir.cpp:126 vs :131 -- not reproduced.)
  Status uploadTaps(const float* taps) noexcept {
    const size_t bytes = mT * (mTapType == SampleType_Float ? 4 : 8);
    UNWRAP_OR_FWD_STATUS(mTaps, mParts.allocator.get()->allocate(bytes));
    CUDA_DEV_PUSH_POP_OR_RET_STATUS(mParts.device());
    // staged through the stream so that `taps` may be freed as soon as createFir returns
    SAFE_CUDA_OR_RET_STATUS(cudaMemcpyAsync(mTaps.get()->data(), taps, bytes, cudaMemcpyHostToDevice, mParts.stream()));
    SAFE_CUDA_OR_RET_STATUS(cudaStreamSynchronize(mParts.stream()));
    return Status_Success;
  }
  // Fir::getNumOutputElements (Fir.cpp:141-187) in its well-defined form; equal to it wherever it does not wrap
  size_t numOutputs() const noexcept {
    const size_t nIn = mParts.port(0).used() / mInBytes;
    return nIn + 1 >= mT ? (nIn + 1 - mT) / mD : 0;
  }

  const SampleType mTapType, mElemType;
  const size_t mD, mT;
  size_t mInBytes = 8, mOutBytes = 8;
  NodeParts mParts;
  Ref<IMemory> mTaps;
  REF_COUNTED(FirFilter);
};

// ---------------------------------------------------------------------------------------------------------------
// CosineSource / ComplexCosineSource: float32 phase bookkeeping exactly as the reference (CosineSource.cpp:51,72,82,
// ComplexCosineSource.cpp:52,72,82); always fills the sink's whole remaining buffer.
// ---------------------------------------------------------------------------------------------------------------
class CosineNode final : public Source {
 public:
  static Result<Source> create(bool complex, float sampleRate, float frequency, ICudaCommandQueue* queue, IFactories* f) noexcept {
    NON_NULL_PARAM_OR_RET(queue);
    CosineNode* node = new (std::nothrow) CosineNode(complex, sampleRate, frequency, queue);
    NON_NULL_OR_RET(node);
    const Status st = node->mParts.init(f, 0);
    if (st != Status_Success) {
      node->unref();
      return ERR_RESULT(st);
    }
    return makeRefResultNonNull<Source>(node);
  }
  size_t getOutputDataSize(size_t port) noexcept final { return port == 0 ? SIZE_MAX : 0; }
  GS_SOURCE_COMMON(mParts, 32 * mElemBytes)
  Status readOutput(IBuffer** bufs, size_t) noexcept final {
    GS_REQUIRE_OUTPUT(bufs);
    IBuffer* out = bufs[0];
    const size_t n = out->range()->remaining() / mElemBytes;
    if (n == 0) return Status_Success;
    const float phiEnd = mPhi + static_cast<float>(n) * mDelta;
    if (mComplex) {
      SAFE_CUDA_OR_RET_STATUS(gsdrCosineC(mPhi, phiEnd, out->writePtr<cuComplex>(), n, mParts.device(), mParts.stream()));
    } else {
      SAFE_CUDA_OR_RET_STATUS(gsdrCosineF(mPhi, phiEnd, out->writePtr<float>(), n, mParts.device(), mParts.stream()));
    }
    mPhi = fmodf(phiEnd, 2.0f * static_cast<float>(M_PI));
    return out->range()->increaseEndOffset(n * mElemBytes);
  }

 private:
  CosineNode(bool complex, float sampleRate, float frequency, ICudaCommandQueue* queue) noexcept
      : mComplex(complex), mElemBytes(complex ? 8 : 4), mDelta(static_cast<float>(2.0 * M_PI * frequency / sampleRate)), mParts(queue) {}
  const bool mComplex;
  const size_t mElemBytes;
  const float mDelta;
  float mPhi = 0.0f;
  NodeParts mParts;
  REF_COUNTED(CosineNode);
};

// ---------------------------------------------------------------------------------------------------------------
// CudaMemcpyFilter: staging node; the input buffer is pinned host memory when the source side is the host
// (reference CudaMemcpyFilter.cpp:32-104)
// ---------------------------------------------------------------------------------------------------------------
class MemcpyFilter final : public Filter {
 public:
  static Result<Filter> create(cudaMemcpyKind kind, ICudaCommandQueue* queue, IFactories* f) noexcept {
    NON_NULL_PARAM_OR_RET(queue);
    MemcpyFilter* node = new (std::nothrow) MemcpyFilter(queue);
    NON_NULL_OR_RET(node);
    const bool hostInput = kind == cudaMemcpyHostToDevice || kind == cudaMemcpyHostToHost;
    Status st = node->mParts.init(f, 1, hostInput);
    if (st == Status_Success) {
      Result<IBufferCopier> copier = f->getCudaBufferCopierFactory()->createBufferCopier(queue, kind);
      st = copier.status;
      node->mCopier = copier.value;
    }
    if (st != Status_Success) {
      node->unref();
      return ERR_RESULT(st);
    }
    return makeRefResultNonNull<Filter>(node);
  }
  GS_SINK_METHODS(mParts, kStep)
  GS_SOURCE_COMMON(mParts, 1)
  size_t getOutputDataSize(size_t port) noexcept final { return port == 0 ? mParts.port(0).used() : 0; }
  Status readOutput(IBuffer** bufs, size_t) noexcept final {
    GS_REQUIRE_OUTPUT(bufs);
    IBuffer* out = bufs[0];
    size_t n = mParts.port(0).used();
    if (n > out->range()->remaining()) n = out->range()->remaining();
    if (n == 0) return Status_Success;
    FWD_IF_ERR(mCopier.get()->copy(out->writePtr(), mParts.port(0).data(), n));
    FWD_IF_ERR(out->range()->increaseEndOffset(n));
    mParts.port(0).consume(n);
    return mParts.port(0).fenceDrained();  // pinned input: the host must not refill it before this copy has run
  }

 private:
  explicit MemcpyFilter(ICudaCommandQueue* queue) noexcept : mParts(queue) {}
  NodeParts mParts;
  Ref<IBufferCopier> mCopier;
  REF_COUNTED(MemcpyFilter);
};

// ---------------------------------------------------------------------------------------------------------------
// FileReader: host Source that fills the caller's buffer with the next bytes of a file (reference FileReader.cpp:48-67)
// ---------------------------------------------------------------------------------------------------------------
class FileReaderNode final : public Source {
 public:
  static Result<Source> create(const char* fileName) noexcept {
    NON_NULL_PARAM_OR_RET(fileName);
    FILE* file = fopen(fileName, "rb");
    if (file == nullptr) {
      gsloge("Cannot open [%s]", fileName);
      return ERR_RESULT(Status_NotFound);
    }
    return makeRefResultNonNull<Source>(new (std::nothrow) FileReaderNode(file));
  }
  size_t getOutputDataSize(size_t port) noexcept final { return port == 0 ? size_t(1) << 16 : 0; }
  size_t getOutputSizeAlignment(size_t) noexcept final { return 1; }
  IBufferCopier* getOutputCopier(size_t) noexcept final { return mCopier.get().get(); }
  Status readOutput(IBuffer** bufs, size_t) noexcept final {
    GS_REQUIRE_OUTPUT(bufs);
    IBuffer* out = bufs[0];
    size_t want = out->range()->remaining();
    if (want > (size_t(1) << 16)) want = size_t(1) << 16;
    const size_t got = fread(out->writePtr(), 1, want, mFile);
    return out->range()->increaseEndOffset(got);
  }

 private:
  explicit FileReaderNode(FILE* file) noexcept : mFile(file), mCopier(newSysMemCopier()) {}
  ~FileReaderNode() final { fclose(mFile); }
  FILE* const mFile;
  Ref<IBufferCopier> mCopier;
  REF_COUNTED_NO_DESTRUCTOR(FileReaderNode);
};

// ---------------------------------------------------------------------------------------------------------------
// Port remapping adapters and the byte-count monitor: pure delegation (reference PortRemappingSink.cpp:19-56,
// PortRemappingSource.cpp:84-125, ReadByteCountMonitor.cpp:44-63)
// ---------------------------------------------------------------------------------------------------------------
class RemapSink final : public IPortRemappingSink {
 public:
  RemapSink() noexcept = default;
  void addPortMapping(size_t outerPort, Sink* innerSink, size_t innerPort) noexcept final {
    try {
      mMap[outerPort] = {Ref<Sink>(innerSink), innerPort};
    } catch (...) {
    }
  }
  Result<IBuffer> requestBuffer(size_t port, size_t byteCount) noexcept final {
    const auto it = mMap.find(port);
    GS_REQUIRE_OR_RET_RESULT_FMT(it != mMap.end(), "Input port [%zu] is not mapped", port);
    return it->second.first.get()->requestBuffer(it->second.second, byteCount);
  }
  Status commitBuffer(size_t port, size_t byteCount) noexcept final {
    const auto it = mMap.find(port);
    GS_REQUIRE_OR_RET_STATUS_FMT(it != mMap.end(), "Input port [%zu] is not mapped", port);
    return it->second.first.get()->commitBuffer(it->second.second, byteCount);
  }
  size_t preferredInputBufferSize(size_t port) noexcept final {
    const auto it = mMap.find(port);
    return it == mMap.end() ? 0 : it->second.first.get()->preferredInputBufferSize(it->second.second);
  }

 private:
  std::map<size_t, std::pair<Ref<Sink>, size_t>> mMap;
  REF_COUNTED(RemapSink);
};

class RemapSource final : public IPortRemappingSource {
 public:
  RemapSource() noexcept = default;
  void addPortMapping(size_t outerPort, Source* innerSource, size_t innerPort) noexcept final {
    try {
      mMap[outerPort] = {Ref<Source>(innerSource), innerPort};
    } catch (...) {
    }
  }
  size_t getOutputDataSize(size_t port) noexcept final {
    const auto it = mMap.find(port);
    return it == mMap.end() ? 0 : it->second.first.get()->getOutputDataSize(it->second.second);
  }
  size_t getOutputSizeAlignment(size_t port) noexcept final {
    const auto it = mMap.find(port);
    return it == mMap.end() ? 1 : it->second.first.get()->getOutputSizeAlignment(it->second.second);
  }
  IBufferCopier* getOutputCopier(size_t port) noexcept final {
    const auto it = mMap.find(port);
    return it == mMap.end() ? nullptr : it->second.first.get()->getOutputCopier(it->second.second);
  }
  // every mapped inner source is read once, with the outer buffers placed at its inner port numbers
  Status readOutput(IBuffer** bufs, size_t numPorts) noexcept final {
    try {
      std::map<Source*, std::vector<IBuffer*>> perSource;
      for (const auto& m : mMap) {
        if (m.first >= numPorts || bufs[m.first] == nullptr) continue;
        auto& v = perSource[m.second.first.get().get()];
        if (v.size() <= m.second.second) v.resize(m.second.second + 1, nullptr);
        v[m.second.second] = bufs[m.first];
      }
      for (auto& s : perSource) FWD_IF_ERR(s.first->readOutput(s.second.data(), s.second.size()));
      return Status_Success;
    }
    IF_CATCH_RETURN_STATUS
  }

 private:
  std::map<size_t, std::pair<Ref<Source>, size_t>> mMap;
  REF_COUNTED(RemapSource);
};

class ByteCountMonitor final : public IReadByteCountMonitor {
 public:
  explicit ByteCountMonitor(Filter* inner) noexcept : mInner(inner) {}
  size_t getByteCountRead(size_t port) noexcept final { return port < mCounts.size() ? mCounts[port] : 0; }
  Result<IBuffer> requestBuffer(size_t port, size_t n) noexcept final { return mInner->requestBuffer(port, n); }
  Status commitBuffer(size_t port, size_t n) noexcept final { return mInner->commitBuffer(port, n); }
  size_t preferredInputBufferSize(size_t port) noexcept final { return mInner->preferredInputBufferSize(port); }
  size_t getOutputDataSize(size_t port) noexcept final { return mInner->getOutputDataSize(port); }
  size_t getOutputSizeAlignment(size_t port) noexcept final { return mInner->getOutputSizeAlignment(port); }
  IBufferCopier* getOutputCopier(size_t port) noexcept final { return mInner->getOutputCopier(port); }
  Status readOutput(IBuffer** bufs, size_t numPorts) noexcept final {
    try {
      std::vector<size_t> before(numPorts, 0);
      for (size_t p = 0; p < numPorts; p++) before[p] = bufs[p] ? bufs[p]->range()->endOffset() : 0;
      FWD_IF_ERR(mInner->readOutput(bufs, numPorts));
      if (mCounts.size() < numPorts) mCounts.resize(numPorts, 0);
      for (size_t p = 0; p < numPorts; p++)
        if (bufs[p]) mCounts[p] += bufs[p]->range()->endOffset() - before[p];
      return Status_Success;
    }
    IF_CATCH_RETURN_STATUS
  }

 private:
  ConstRef<Filter> mInner;
  std::vector<size_t> mCounts;
  REF_COUNTED(ByteCountMonitor);
};

// ---------------------------------------------------------------------------------------------------------------
// Factories (typed create + JSON create).  JSON keys are the reference's: factories/FirFactory.h:32-49,
// CosineSourceFactory.h:34-46, QuadDemodFactory.h:35-70, the others take {"commandQueue": id}.
// ---------------------------------------------------------------------------------------------------------------
Result<SampleType> parseSampleType(const Json& v) {
  const std::string& s = v.str();
  if (s == "FloatComplex" || s == "floatComplex" || s == "ComplexFloat" || s == "complex") return makeValResult<SampleType>(SampleType_FloatComplex);
  if (s == "Float" || s == "float") return makeValResult<SampleType>(SampleType_Float);
  if (s == "Int8Complex" || s == "int8Complex") return makeValResult<SampleType>(SampleType_Int8Complex);
  gsloge("Unknown sample type [%s]", s.c_str());
  return ERR_RESULT(Status_ParseError);
}

// "commandQueue" is the key the reference's node factories read (factories/FirFactory.h:32-49 etc.); "commandQueueId" is
// the key its RfToPcmAudioFactory.cpp:223-296 writes into the Component it builds -- both are accepted
Result<ICudaCommandQueue> queueFromJson(IFactories* f, const Json& params) {
  const char* key = params.contains("commandQueue") ? "commandQueue" : "commandQueueId";
  return f->getCommandQueueFactory()->getCudaCommandQueue(params.at(key).str().c_str());
}

#define GS_JSON_CREATE_BEGIN                              \
  Result<Node> create(const char* json) noexcept final {  \
    try {                                                 \
      const Json params = Json::parse(json);              \
      Ref<ICudaCommandQueue> queue;                       \
      UNWRAP_OR_FWD_RESULT(queue, queueFromJson(mFactories, params));
#define GS_JSON_CREATE_END                                \
    } catch (const std::invalid_argument& e) {            \
      gsloge("Bad node parameters: %s", e.what());        \
      return ERR_RESULT(Status_ParseError);               \
    }                                                     \
    IF_CATCH_RETURN_RESULT                                \
  }

class MapFactory final : public ICudaFilterFactory {
 public:
  MapFactory(IFactories* f, MapKind kind) noexcept : mFactories(f), mKind(kind) {}
  Result<Filter> createFilter(ICudaCommandQueue* queue) noexcept final { return MapFilter::create(mKind, 0.0f, queue, mFactories); }
  GS_JSON_CREATE_BEGIN
  return ResultCast<Node>(createFilter(queue.get()));
  GS_JSON_CREATE_END
 private:
  IFactories* const mFactories;
  const MapKind mKind;
  REF_COUNTED(MapFactory);
};

class MultiplyFactory final : public ICudaFilterFactory {
 public:
  explicit MultiplyFactory(IFactories* f) noexcept : mFactories(f) {}
  Result<Filter> createFilter(ICudaCommandQueue* queue) noexcept final { return MultiplyFilter::create(queue, mFactories); }
  GS_JSON_CREATE_BEGIN
  return ResultCast<Node>(createFilter(queue.get()));
  GS_JSON_CREATE_END
 private:
  IFactories* const mFactories;
  REF_COUNTED(MultiplyFactory);
};

class AddConstFactory final : public IAddConstFactory {
 public:
  explicit AddConstFactory(IFactories* f) noexcept : mFactories(f) {}
  Result<Filter> createAddConst(float c, ICudaCommandQueue* queue) noexcept final { return MapFilter::create(MapKind::AddConst, c, queue, mFactories); }
  GS_JSON_CREATE_BEGIN
  return ResultCast<Node>(createAddConst(static_cast<float>(params.at("addValueToAmplitude").num()), queue.get()));
  GS_JSON_CREATE_END
 private:
  IFactories* const mFactories;
  REF_COUNTED(AddConstFactory);
};

class AddToMagnitudeFactory final : public IAddConstToVectorLengthFactory {
 public:
  explicit AddToMagnitudeFactory(IFactories* f) noexcept : mFactories(f) {}
  Result<Filter> createAddConstToVectorLength(float c, ICudaCommandQueue* queue) noexcept final {
    return MapFilter::create(MapKind::AddToMagnitude, c, queue, mFactories);
  }
  GS_JSON_CREATE_BEGIN
  return ResultCast<Node>(createAddConstToVectorLength(static_cast<float>(params.at("addValueToMagnitude").num()), queue.get()));
  GS_JSON_CREATE_END
 private:
  IFactories* const mFactories;
  REF_COUNTED(AddToMagnitudeFactory);
};

class CosineFactory final : public ICosineSourceFactory {
 public:
  explicit CosineFactory(IFactories* f) noexcept : mFactories(f) {}
  Result<Source> createCosineSource(SampleType type, float sampleRate, float frequency, ICudaCommandQueue* queue) noexcept final {
    GS_REQUIRE_OR_RET_RESULT_FMT(type == SampleType_Float || type == SampleType_FloatComplex, "Unsupported cosine sample type [%u]", type);
    return CosineNode::create(type == SampleType_FloatComplex, sampleRate, frequency, queue, mFactories);
  }
  GS_JSON_CREATE_BEGIN
  SampleType type;
  UNWRAP_OR_FWD_RESULT(type, parseSampleType(params.at("sampleType")));
  return ResultCast<Node>(createCosineSource(type, static_cast<float>(params.at("sampleRate").num()),
                                             static_cast<float>(params.at("frequency").num()), queue.get()));
  GS_JSON_CREATE_END
 private:
  IFactories* const mFactories;
  REF_COUNTED(CosineFactory);
};

class FirFactory final : public IFirFactory {
 public:
  explicit FirFactory(IFactories* f) noexcept : mFactories(f) {}
  Result<Filter> createFir(SampleType tapType, SampleType elementType, size_t decimation, const float* taps, size_t tapCount,
                           ICudaCommandQueue* queue) noexcept final {
    return FirFilter::create(tapType, elementType, decimation, taps, tapCount, queue, mFactories);
  }
  GS_JSON_CREATE_BEGIN
  SampleType tapType, elementType;
  UNWRAP_OR_FWD_RESULT(tapType, parseSampleType(params.at("tapType")));
  // "elementType" (FirFactory.h:40) or "signalType" (what RfToPcmAudioFactory.cpp:252,287 emits)
  UNWRAP_OR_FWD_RESULT(elementType, parseSampleType(params.at(params.contains("elementType") ? "elementType" : "signalType")));
  std::vector<float> taps;
  for (const Json& t : params.at("taps").array()) taps.push_back(static_cast<float>(t.num()));
  // complex taps arrive as (re, im) pairs: tapCount counts taps, not floats.  (The reference passes taps.size() floats
  // as the tap count, FirFactory.h:44-48, and would read past the list; an odd-length list is rejected here.)
  GS_REQUIRE_OR_RET_RESULT(tapType == SampleType_Float || taps.size() % 2 == 0, "Complex taps must be given as (re, im) pairs");
  const size_t tapCount = tapType == SampleType_Float ? taps.size() : taps.size() / 2;
  return ResultCast<Node>(createFir(tapType, elementType, static_cast<size_t>(params.at("decimation").num()), taps.data(), tapCount, queue.get()));
  GS_JSON_CREATE_END
 private:
  IFactories* const mFactories;
  REF_COUNTED(FirFactory);
};

class QuadDemodFactory final : public IQuadDemodFactory {
 public:
  explicit QuadDemodFactory(IFactories* f) noexcept : mFactories(f) {}
  Result<Filter> createQuadDemod(Modulation modulation, float rfSampleRate, float fskDeviation, ICudaCommandQueue* queue) noexcept final {
    switch (modulation) {
      case Modulation_Am: return MapFilter::create(MapKind::QuadAm, 0.0f, queue, mFactories);
      case Modulation_Fm: {
        const float gain = rfSampleRate / (2.0f * static_cast<float>(M_PI) * fskDeviation * 5);  // QuadDemodFactory.h:108-110
        return MapFilter::create(MapKind::QuadFm, gain, queue, mFactories);
      }
      default: gsloge("Modulation [%u] is not supported", modulation); return ERR_RESULT(Status_InvalidArgument);
    }
  }
  GS_JSON_CREATE_BEGIN
  // a string (QuadDemodFactory.h:35-70) or the Modulation enum value (what RfToPcmAudioFactory.cpp:266 writes)
  const Json& mj = params.at("modulation");
  bool fm;
  if (mj.kind == Json::Number) {
    GS_REQUIRE_OR_RET_RESULT_FMT(mj.num() == Modulation_Am || mj.num() == Modulation_Fm, "Unknown modulation [%g]", mj.num());
    fm = mj.num() == Modulation_Fm;
  } else {
    const std::string& m = mj.str();
    GS_REQUIRE_OR_RET_RESULT_FMT(m == "AM" || m == "FM" || m == "am" || m == "fm", "Unknown modulation [%s]", m.c_str());
    fm = m == "FM" || m == "fm";
  }
  const float rate = fm ? static_cast<float>(params.at("sampleRate").num()) : 0.0f;
  const float dev = fm ? static_cast<float>(params.at("fskDeviation").num()) : 0.0f;
  return ResultCast<Node>(createQuadDemod(fm ? Modulation_Fm : Modulation_Am, rate, dev, queue.get()));
  GS_JSON_CREATE_END
 private:
  IFactories* const mFactories;
  REF_COUNTED(QuadDemodFactory);
};

class MemcpyFactory final : public ICudaMemcpyFilterFactory {
 public:
  explicit MemcpyFactory(IFactories* f) noexcept : mFactories(f) {}
  Result<Filter> createCudaMemcpy(cudaMemcpyKind kind, ICudaCommandQueue* queue) noexcept final { return MemcpyFilter::create(kind, queue, mFactories); }
  GS_JSON_CREATE_BEGIN
  const std::string& k = params.at("memcpyKind").str();
  cudaMemcpyKind kind;
  if (k == "hostToDevice") kind = cudaMemcpyHostToDevice;
  else if (k == "deviceToHost") kind = cudaMemcpyDeviceToHost;
  else if (k == "deviceToDevice") kind = cudaMemcpyDeviceToDevice;
  else if (k == "hostToHost") kind = cudaMemcpyHostToHost;
  else return ERR_RESULT(Status_ParseError);
  return ResultCast<Node>(createCudaMemcpy(kind, queue.get()));
  GS_JSON_CREATE_END
 private:
  IFactories* const mFactories;
  REF_COUNTED(MemcpyFactory);
};

class FileReaderFactory final : public IFileReaderFactory {
 public:
  FileReaderFactory() noexcept = default;
  Result<Source> createFileReader(const char* fileName) noexcept final { return FileReaderNode::create(fileName); }
  Result<Node> create(const char* json) noexcept final {
    try {
      return ResultCast<Node>(createFileReader(Json::parse(json).at("fileName").str().c_str()));
    } catch (const std::invalid_argument&) {
      return ERR_RESULT(Status_ParseError);
    }
    IF_CATCH_RETURN_RESULT
  }
  REF_COUNTED(FileReaderFactory);
};

// Hardware source and audio-codec sink: outside this library's scope (libhackrf / FFmpeg); the vtable slots exist so
// that IFactories keeps the reference's layout, and every call answers Status_NotFound.
class NoHackrfFactory final : public IHackrfSourceFactory {
 public:
  NoHackrfFactory() noexcept = default;
  Result<Node> create(const char*) noexcept final { return ERR_RESULT(Status_NotFound); }
  Result<IHackrfSource> createHackrfSource(int32_t, uint64_t, double, size_t) noexcept final {
    gsloge("HackRF capture is not part of this library: feed int8 IQ through a CudaMemcpyFilter or the fused chain node");
    return ERR_RESULT(Status_NotFound);
  }
  REF_COUNTED(NoHackrfFactory);
};
class NoAacFactory final : public IAacFileWriterFactory {
 public:
  NoAacFactory() noexcept = default;
  Result<Node> create(const char*) noexcept final { return ERR_RESULT(Status_NotFound); }
  Result<Sink> createAacFileWriter(const char*, int32_t, int32_t, ICudaCommandQueue*) noexcept final {
    gsloge("AAC encoding is not part of this library");
    return ERR_RESULT(Status_NotFound);
  }
  REF_COUNTED(NoAacFactory);
};

class RemapSinkFactory final : public IPortRemappingSinkFactory {
 public:
  RemapSinkFactory() noexcept = default;
  Result<IPortRemappingSink> create() noexcept final { return makeRefResultNonNull<IPortRemappingSink>(new (std::nothrow) RemapSink()); }
  REF_COUNTED(RemapSinkFactory);
};
class RemapSourceFactory final : public IPortRemappingSourceFactory {
 public:
  RemapSourceFactory() noexcept = default;
  Result<IPortRemappingSource> create() noexcept final { return makeRefResultNonNull<IPortRemappingSource>(new (std::nothrow) RemapSource()); }
  REF_COUNTED(RemapSourceFactory);
};
class MonitorFactory final : public IReadByteCountMonitorFactory {
 public:
  MonitorFactory() noexcept = default;
  Result<IReadByteCountMonitor> create(Filter* monitored) noexcept final {
    NON_NULL_PARAM_OR_RET(monitored);
    return makeRefResultNonNull<IReadByteCountMonitor>(new (std::nothrow) ByteCountMonitor(monitored));
  }
  REF_COUNTED(MonitorFactory);
};

}  // namespace

ICudaMemcpyFilterFactory* newCudaMemcpyFilterFactory(IFactories* f) noexcept { return new (std::nothrow) MemcpyFactory(f); }
IAacFileWriterFactory* newAacFileWriterFactory() noexcept { return new (std::nothrow) NoAacFactory(); }
IAddConstFactory* newAddConstFactory(IFactories* f) noexcept { return new (std::nothrow) AddConstFactory(f); }
IAddConstToVectorLengthFactory* newAddConstToVectorLengthFactory(IFactories* f) noexcept { return new (std::nothrow) AddToMagnitudeFactory(f); }
ICosineSourceFactory* newCosineSourceFactory(IFactories* f) noexcept { return new (std::nothrow) CosineFactory(f); }
IFileReaderFactory* newFileReaderFactory() noexcept { return new (std::nothrow) FileReaderFactory(); }
IFirFactory* newFirFactory(IFactories* f) noexcept { return new (std::nothrow) FirFactory(f); }
IHackrfSourceFactory* newHackrfSourceFactory() noexcept { return new (std::nothrow) NoHackrfFactory(); }
ICudaFilterFactory* newInt8ToFloatFactory(IFactories* f) noexcept { return new (std::nothrow) MapFactory(f, MapKind::Int8ToFloat); }
ICudaFilterFactory* newMagnitudeFactory(IFactories* f) noexcept { return new (std::nothrow) MapFactory(f, MapKind::Magnitude); }
ICudaFilterFactory* newMultiplyFactory(IFactories* f) noexcept { return new (std::nothrow) MultiplyFactory(f); }
IQuadDemodFactory* newQuadDemodFactory(IFactories* f) noexcept { return new (std::nothrow) QuadDemodFactory(f); }
IPortRemappingSinkFactory* newPortRemappingSinkFactory() noexcept { return new (std::nothrow) RemapSinkFactory(); }
IPortRemappingSourceFactory* newPortRemappingSourceFactory() noexcept { return new (std::nothrow) RemapSourceFactory(); }
IReadByteCountMonitorFactory* newReadByteCountMonitorFactory() noexcept { return new (std::nothrow) MonitorFactory(); }

}  // namespace gs
