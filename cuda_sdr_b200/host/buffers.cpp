// Memory, ranges, buffers, slices, copiers, memset and pools behind the buffer interfaces of
// include/gpusdrpipeline/abi/buffers.h.
//
// What is different from the reference's implementation (src/buffers/*.cpp) while keeping its interfaces:
//   * RelocatableBuffer::relocate() moves data INSIDE its one allocation (a stream-ordered copy, chunked when source
//     and destination overlap) instead of copying into a second, equally large "twin" allocation and swapping
//     (reference RelocatableResizableBuffer.cpp:79-103): half the memory and no second allocation to keep warm.
//   * The per-port input buffers of this library's own nodes do not use it at all: see PortInput in nodes.cpp, where
//     consuming input is pointer arithmetic and compaction happens only when the tail runs out of room.
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>

#include "internal.h"

namespace gs {
namespace {

// ---------------------------------------------------------------------------------------------------------------
class Range final : public IBufferRangeMutableCapacity {
 public:
  size_t capacity() const noexcept final { return mCapacity; }
  size_t offset() const noexcept final { return mOffset; }
  size_t endOffset() const noexcept final { return mEnd; }
  Status setUsedRange(size_t offset, size_t endOffset) noexcept final {
    GS_REQUIRE_OR_RET_STATUS_FMT(offset <= endOffset, "used range: offset [%zu] > end offset [%zu]", offset, endOffset);
    GS_REQUIRE_OR_RET_STATUS_FMT(endOffset <= mCapacity, "used range: end offset [%zu] > capacity [%zu]", endOffset, mCapacity);
    mOffset = offset;
    mEnd = endOffset;
    return Status_Success;
  }
  void setCapacity(size_t capacity) noexcept final {
    mCapacity = capacity;
    if (mEnd > capacity) mEnd = capacity;
    if (mOffset > mEnd) mOffset = mEnd;
  }

 private:
  size_t mCapacity = 0, mOffset = 0, mEnd = 0;
  REF_COUNTED(Range);

 public:
  Range() noexcept = default;
};

class RangeFactory final : public IBufferRangeFactory {
 public:
  RangeFactory() noexcept = default;
  Result<IBufferRangeMutableCapacity> createBufferRange() const noexcept final { return newBufferRange(); }
  REF_COUNTED(RangeFactory);
};

// ---------------------------------------------------------------------------------------------------------------
class SysMemory final : public IMemory {
 public:
  SysMemory(void* raw, size_t capacity) noexcept : mData(static_cast<uint8_t*>(raw)), mCapacity(capacity) {}
  uint8_t* data() noexcept final { return mData; }
  const uint8_t* data() const noexcept final { return mData; }
  size_t capacity() const noexcept final { return mCapacity; }

 private:
  uint8_t* const mData;
  const size_t mCapacity;
  ~SysMemory() final { free(mData); }
  REF_COUNTED_NO_DESTRUCTOR(SysMemory);
};

class SysMemAllocator final : public IAllocator {
 public:
  SysMemAllocator() noexcept = default;
  Result<IMemory> allocate(size_t size) noexcept final {
    void* raw = aligned_alloc(64, roundUp(size == 0 ? 64 : size, 64));
    if (raw == nullptr) return ERR_RESULT(Status_OutOfMemory);
    return makeRefResultNonNull<IMemory>(new (std::nothrow) SysMemory(raw, size));
  }
  REF_COUNTED(SysMemAllocator);
};

// Device memory from the stream-ordered pool of the queue's device (freed on the same stream, so a buffer may be
// released while work that uses it is still queued), or pinned host memory.  Reference: CudaAllocator.cpp:32-110.
class CudaMemory final : public IMemory {
 public:
  CudaMemory(void* raw, size_t capacity, bool host, ICudaCommandQueue* queue) noexcept
      : mData(static_cast<uint8_t*>(raw)), mCapacity(capacity), mHost(host), mQueue(queue) {}
  uint8_t* data() noexcept final { return mData; }
  const uint8_t* data() const noexcept final { return mData; }
  size_t capacity() const noexcept final { return mCapacity; }

 private:
  uint8_t* const mData;
  const size_t mCapacity;
  const bool mHost;
  ConstRef<ICudaCommandQueue> mQueue;
  ~CudaMemory() final {
    CudaDevicePushPop device(mQueue->cudaDevice());
    if (mHost) {
      SAFE_CUDA_WARN_ONLY(cudaStreamSynchronize(mQueue->cudaStream()));
      SAFE_CUDA_WARN_ONLY(cudaFreeHost(mData));
    } else {
      SAFE_CUDA_WARN_ONLY(cudaFreeAsync(mData, mQueue->cudaStream()));
    }
  }
  REF_COUNTED_NO_DESTRUCTOR(CudaMemory);
};

class CudaAllocator final : public IAllocator {
 public:
  CudaAllocator(ICudaCommandQueue* queue, size_t alignment, bool host) noexcept
      : mQueue(queue), mAlignment(alignment == 0 ? 1 : alignment), mHost(host) {}
  Result<IMemory> allocate(size_t size) noexcept final {
    // capacity is rounded up to the alignment (callers size launches from it, tests/CosineSourceTests.cpp:31-33);
    // cudaMalloc* and cudaHostAlloc return pointers aligned to at least 256 bytes
    GS_REQUIRE_OR_RET_RESULT(mAlignment <= 256, "CUDA allocations are aligned to 256 bytes at most");
    const size_t bytes = roundUp(size == 0 ? mAlignment : size, mAlignment);
    CUDA_DEV_PUSH_POP_OR_RET_RESULT(mQueue->cudaDevice());
    void* raw = nullptr;
    if (mHost) {
      SAFE_CUDA_OR_RET_RESULT(cudaHostAlloc(&raw, bytes, cudaHostAllocDefault));
    } else {
      SAFE_CUDA_OR_RET_RESULT(cudaMallocAsync(&raw, bytes, mQueue->cudaStream()));
    }
    IMemory* memory = new (std::nothrow) CudaMemory(raw, bytes, mHost, mQueue);
    if (memory == nullptr) {
      if (mHost) cudaFreeHost(raw); else cudaFreeAsync(raw, mQueue->cudaStream());
    }
    return makeRefResultNonNull(memory);
  }

 private:
  ConstRef<ICudaCommandQueue> mQueue;
  const size_t mAlignment;
  const bool mHost;
  REF_COUNTED(CudaAllocator);
};

class CudaAllocatorFactory final : public ICudaAllocatorFactory {
 public:
  CudaAllocatorFactory() noexcept = default;
  Result<IAllocator> createCudaAllocator(ICudaCommandQueue* queue, size_t alignment, bool useHostMemory) noexcept final {
    NON_NULL_PARAM_OR_RET(queue);
    return makeRefResultNonNull<IAllocator>(new (std::nothrow) CudaAllocator(queue, alignment, useHostMemory));
  }
  REF_COUNTED(CudaAllocatorFactory);
};

// ---------------------------------------------------------------------------------------------------------------
class SysMemCopier final : public IBufferCopier {
 public:
  SysMemCopier() noexcept = default;
  Status copy(void* dst, const void* src, size_t length) const noexcept final {
    if (length) memmove(dst, src, length);
    return Status_Success;
  }
  REF_COUNTED(SysMemCopier);
};

class CudaCopier final : public IBufferCopier {
 public:
  CudaCopier(ICudaCommandQueue* queue, cudaMemcpyKind kind) noexcept : mQueue(queue), mKind(kind) {}
  Status copy(void* dst, const void* src, size_t length) const noexcept final {
    if (length == 0) return Status_Success;
    CUDA_DEV_PUSH_POP_OR_RET_STATUS(mQueue->cudaDevice());
    SAFE_CUDA_OR_RET_STATUS(cudaMemcpyAsync(dst, src, length, mKind, mQueue->cudaStream()));
    return Status_Success;
  }

 private:
  ConstRef<ICudaCommandQueue> mQueue;
  const cudaMemcpyKind mKind;
  REF_COUNTED(CudaCopier);
};

class CudaCopierFactory final : public ICudaBufferCopierFactory {
 public:
  CudaCopierFactory() noexcept = default;
  Result<IBufferCopier> createBufferCopier(ICudaCommandQueue* queue, cudaMemcpyKind kind) noexcept final {
    NON_NULL_PARAM_OR_RET(queue);
    return makeRefResultNonNull<IBufferCopier>(new (std::nothrow) CudaCopier(queue, kind));
  }
  REF_COUNTED(CudaCopierFactory);
};

class SysMemSet final : public IMemSet {
 public:
  SysMemSet() noexcept = default;
  Status memSet(void* data, uint8_t value, size_t byteCount) noexcept final {
    if (byteCount) memset(data, value, byteCount);
    return Status_Success;
  }
  REF_COUNTED(SysMemSet);
};

class CudaMemSet final : public IMemSet {
 public:
  explicit CudaMemSet(ICudaCommandQueue* queue) noexcept : mQueue(queue) {}
  Status memSet(void* data, uint8_t value, size_t byteCount) noexcept final {
    if (byteCount == 0) return Status_Success;
    CUDA_DEV_PUSH_POP_OR_RET_STATUS(mQueue->cudaDevice());
    SAFE_CUDA_OR_RET_STATUS(cudaMemsetAsync(data, value, byteCount, mQueue->cudaStream()));
    return Status_Success;
  }

 private:
  ConstRef<ICudaCommandQueue> mQueue;
  REF_COUNTED(CudaMemSet);
};

class CudaMemSetFactory final : public ICudaMemSetFactory {
 public:
  CudaMemSetFactory() noexcept = default;
  Result<IMemSet> create(ICudaCommandQueue* queue) noexcept final {
    NON_NULL_PARAM_OR_RET(queue);
    return makeRefResultNonNull<IMemSet>(new (std::nothrow) CudaMemSet(queue));
  }
  REF_COUNTED(CudaMemSetFactory);
};

// ---------------------------------------------------------------------------------------------------------------
class OwnedBuffer final : public IBuffer {
 public:
  OwnedBuffer(IMemory* memory, IBufferRangeMutableCapacity* range) noexcept : mMemory(memory), mRange(range) {}
  uint8_t* base() noexcept final { return mMemory->data(); }
  const uint8_t* base() const noexcept final { return mMemory->data(); }
  IBufferRange* range() noexcept final { return mRange.get(); }
  const IBufferRange* range() const noexcept final { return mRange.get(); }

 private:
  ConstRef<IMemory> mMemory;
  ConstRef<IBufferRangeMutableCapacity> mRange;
  REF_COUNTED(OwnedBuffer);
};

// A window [start, end) of another buffer with its own used range, initialised to the part of the parent's used range
// that falls inside the window (reference BufferSlice.cpp:103-186).  Keeps the parent alive.
class SliceBuffer final : public IBuffer {
 public:
  SliceBuffer(IBuffer* parent, size_t start, IBufferRangeMutableCapacity* range) noexcept : mParent(parent), mStart(start), mRange(range) {}
  uint8_t* base() noexcept final { return mParent->base() + mStart; }
  const uint8_t* base() const noexcept final { return mParent->base() + mStart; }
  IBufferRange* range() noexcept final { return mRange.get(); }
  const IBufferRange* range() const noexcept final { return mRange.get(); }

 private:
  ConstRef<IBuffer> mParent;
  const size_t mStart;
  ConstRef<IBufferRangeMutableCapacity> mRange;
  REF_COUNTED(SliceBuffer);
};

class SliceFactory final : public IBufferSliceFactory {
 public:
  SliceFactory() noexcept = default;
  Result<IBuffer> slice(IBuffer* buffer, size_t start, size_t end) noexcept final { return newBufferSlice(buffer, start, end); }
  REF_COUNTED(SliceFactory);
};

class BufferFactory final : public IBufferFactory {
 public:
  explicit BufferFactory(IAllocator* allocator) noexcept : mAllocator(allocator) {}
  Result<IBuffer> createBuffer(size_t size) noexcept final { return newOwnedBuffer(mAllocator, size); }

 private:
  ConstRef<IAllocator> mAllocator;
  REF_COUNTED(BufferFactory);
};

// ---------------------------------------------------------------------------------------------------------------
// One allocation that can grow (allocate + copy the used bytes + swap) and move its used bytes around.
class RelocatableBuffer final : public IRelocatableResizableBuffer {
 public:
  RelocatableBuffer(IAllocator* allocator, const IBufferCopier* copier, IMemory* memory, IBufferRangeMutableCapacity* range) noexcept
      : mAllocator(allocator), mCopier(copier), mMemory(memory), mRange(range) {}

  uint8_t* base() noexcept final { return mMemory.get()->data(); }
  const uint8_t* base() const noexcept final { return mMemory.get()->data(); }
  IBufferRange* range() noexcept final { return mRange.get(); }
  const IBufferRange* range() const noexcept final { return mRange.get(); }

  Status resize(size_t newSize) noexcept final {
    if (newSize <= mRange->capacity()) return Status_Success;  // never shrinks (reference ResizableBuffer semantics)
    Ref<IMemory> bigger;
    UNWRAP_OR_FWD_STATUS(bigger, mAllocator->allocate(newSize));
    const size_t offset = mRange->offset(), used = mRange->used();
    FWD_IF_ERR(mCopier->copy(bigger.get()->data() + offset, base() + offset, used));
    mMemory = bigger;
    mRange->setCapacity(bigger.get()->capacity());
    return Status_Success;
  }

  Status relocate(size_t dstOffset, size_t srcOffset, size_t length) noexcept final {
    GS_REQUIRE_OR_RET_STATUS(srcOffset + length <= mRange->capacity() && dstOffset + length <= mRange->capacity(), "relocate out of range");
    if (dstOffset != srcOffset && length > 0) {
      uint8_t* const b = base();
      const size_t gap = dstOffset < srcOffset ? srcOffset - dstOffset : dstOffset - srcOffset;
      if (gap >= length) {
        FWD_IF_ERR(mCopier->copy(b + dstOffset, b + srcOffset, length));
      } else if (dstOffset < srcOffset && gap * 64 >= length) {
        // overlapping move towards the start: stream-ordered chunks of `gap` bytes never overlap themselves
        for (size_t done = 0; done < length; done += gap) {
          const size_t n = length - done < gap ? length - done : gap;
          FWD_IF_ERR(mCopier->copy(b + dstOffset + done, b + srcOffset + done, n));
        }
      } else {
        Ref<IMemory> scratch;  // heavy overlap: bounce through a temporary of the moved size
        UNWRAP_OR_FWD_STATUS(scratch, mAllocator->allocate(length));
        FWD_IF_ERR(mCopier->copy(scratch.get()->data(), b + srcOffset, length));
        FWD_IF_ERR(mCopier->copy(b + dstOffset, scratch.get()->data(), length));
      }
    }
    return mRange->setUsedRange(dstOffset, dstOffset + length);
  }

 private:
  ConstRef<IAllocator> mAllocator;
  ConstRef<const IBufferCopier> mCopier;
  Ref<IMemory> mMemory;
  ConstRef<IBufferRangeMutableCapacity> mRange;
  REF_COUNTED(RelocatableBuffer);
};

Result<IRelocatableResizableBuffer> newRelocatableBuffer(IAllocator* allocator, const IBufferCopier* copier, size_t size) noexcept {
  Ref<IMemory> memory;
  Ref<IBufferRangeMutableCapacity> range;
  UNWRAP_OR_FWD_RESULT(memory, allocator->allocate(size));
  UNWRAP_OR_FWD_RESULT(range, newBufferRange());
  range.get()->setCapacity(memory.get()->capacity());
  return makeRefResultNonNull<IRelocatableResizableBuffer>(
      new (std::nothrow) RelocatableBuffer(allocator, copier, memory.get(), range.get()));
}

// IResizableBuffer view of the same implementation
class ResizableBuffer final : public IResizableBuffer {
 public:
  explicit ResizableBuffer(IRelocatableResizableBuffer* impl) noexcept : mImpl(impl) {}
  uint8_t* base() noexcept final { return mImpl->base(); }
  const uint8_t* base() const noexcept final { return mImpl->base(); }
  IBufferRange* range() noexcept final { return mImpl->range(); }
  const IBufferRange* range() const noexcept final { return mImpl->range(); }
  Status resize(size_t newSize) noexcept final { return mImpl->resize(newSize); }

 private:
  ConstRef<IRelocatableResizableBuffer> mImpl;
  REF_COUNTED(ResizableBuffer);
};

class ResizableFactory final : public IResizableBufferFactory {
 public:
  ResizableFactory(IAllocator* allocator, const IBufferCopier* copier) noexcept : mAllocator(allocator), mCopier(copier) {}
  Result<IResizableBuffer> createResizableBuffer(size_t size) noexcept final {
    Ref<IRelocatableResizableBuffer> impl;
    UNWRAP_OR_FWD_RESULT(impl, newRelocatableBuffer(mAllocator, mCopier, size));
    return makeRefResultNonNull<IResizableBuffer>(new (std::nothrow) ResizableBuffer(impl.get()));
  }

 private:
  ConstRef<IAllocator> mAllocator;
  ConstRef<const IBufferCopier> mCopier;
  REF_COUNTED(ResizableFactory);
};

class RelocatableFactory final : public IRelocatableResizableBufferFactory {
 public:
  RelocatableFactory(IAllocator* allocator, const IBufferCopier* copier) noexcept : mAllocator(allocator), mCopier(copier) {}
  Result<IRelocatableResizableBuffer> createRelocatableBuffer(size_t size) const noexcept final {
    return newRelocatableBuffer(mAllocator, mCopier, size);
  }

 private:
  ConstRef<IAllocator> mAllocator;
  ConstRef<const IBufferCopier> mCopier;
  REF_COUNTED(RelocatableFactory);
};

// ---------------------------------------------------------------------------------------------------------------
class BufferUtil final : public IBufferUtil {
 public:
  BufferUtil() noexcept = default;
  Status appendToBuffer(IBuffer* buffer, const void* src, size_t count, const IBufferCopier* copier) const noexcept final {
    GS_REQUIRE_OR_RET_STATUS(buffer->range()->remaining() >= count, "appendToBuffer: not enough room");
    FWD_IF_ERR(copier->copy(buffer->writePtr(), src, count));
    return buffer->range()->increaseEndOffset(count);
  }
  Status readFromBuffer(void* dst, IBuffer* buffer, size_t count, const IBufferCopier* copier) const noexcept final {
    GS_REQUIRE_OR_RET_STATUS(buffer->range()->used() >= count, "readFromBuffer: not enough data");
    FWD_IF_ERR(copier->copy(dst, buffer->readPtr(), count));
    return buffer->range()->increaseOffset(count);
  }
  Status moveFromBuffer(IBuffer* dst, IBuffer* src, size_t count, const IBufferCopier* copier) const noexcept final {
    GS_REQUIRE_OR_RET_STATUS(dst->range()->remaining() >= count && src->range()->used() >= count, "moveFromBuffer: range too small");
    FWD_IF_ERR(copier->copy(dst->writePtr(), src->readPtr(), count));
    FWD_IF_ERR(dst->range()->increaseEndOffset(count));
    return src->range()->increaseOffset(count);
  }
  REF_COUNTED(BufferUtil);
};

// ---------------------------------------------------------------------------------------------------------------
// At most maxBufferCount buffers of one size; a buffer goes back to the free list when its last reference is dropped
// (reference BufferPool.cpp:54-103).
class BufferPool final : public IBufferPool {
 public:
  BufferPool(size_t maxCount, size_t bufferSize, IBufferFactory* factory) noexcept : mMax(maxCount), mSize(bufferSize), mFactory(factory) {}
  size_t getBufferSize() const noexcept final { return mSize; }
  Result<IBuffer> getBuffer() noexcept final { return take(true); }
  Result<IBuffer> tryGetBuffer() noexcept final { return take(false); }

 private:
  class Loan final : public IBuffer {
   public:
    Loan(BufferPool* pool, IBuffer* inner) noexcept : mPool(pool), mInner(inner) {}
    uint8_t* base() noexcept final { return mInner->base(); }
    const uint8_t* base() const noexcept final { return mInner->base(); }
    IBufferRange* range() noexcept final { return mInner->range(); }
    const IBufferRange* range() const noexcept final { return mInner->range(); }

   private:
    ConstRef<BufferPool> mPool;
    IBuffer* const mInner;  // owned by the pool's list
    ~Loan() final { mPool->giveBack(mInner); }
    REF_COUNTED_NO_DESTRUCTOR(Loan);
  };

  Result<IBuffer> take(bool wait) noexcept {
    std::unique_lock<std::mutex> lock(mMutex);
    for (;;) {
      if (!mFree.empty()) {
        IBuffer* inner = mFree.front();
        mFree.pop_front();
        inner->range()->clearRange();
        return makeRefResultNonNull<IBuffer>(new (std::nothrow) Loan(this, inner));
      }
      if (mCreated < mMax) {
        Result<IBuffer> made = mFactory->createBuffer(mSize);
        if (made.status != Status_Success) return made;
        made.value->ref();  // the pool's own reference, for the pool's lifetime
        mAll.push_back(made.value);
        mCreated++;
        return makeRefResultNonNull<IBuffer>(new (std::nothrow) Loan(this, made.value));
      }
      if (!wait) return makeRefResultNullable<IBuffer>(nullptr);
      mAvailable.wait(lock);
    }
  }
  void giveBack(IBuffer* inner) noexcept {
    {
      std::lock_guard<std::mutex> lock(mMutex);
      mFree.push_back(inner);
    }
    mAvailable.notify_one();
  }

  const size_t mMax, mSize;
  ConstRef<IBufferFactory> mFactory;
  std::mutex mMutex;
  std::condition_variable mAvailable;
  std::deque<IBuffer*> mFree;
  std::vector<IBuffer*> mAll;
  size_t mCreated = 0;
  ~BufferPool() final {
    for (IBuffer* b : mAll) b->unref();
  }
  REF_COUNTED_NO_DESTRUCTOR(BufferPool);
};

class BufferPoolFactory final : public IBufferPoolFactory {
 public:
  BufferPoolFactory(size_t maxCount, IBufferFactory* factory) noexcept : mMax(maxCount), mFactory(factory) {}
  Result<IBufferPool> createBufferPool(size_t bufferSize) noexcept final {
    return makeRefResultNonNull<IBufferPool>(newBufferPool(mMax, bufferSize, mFactory));
  }

 private:
  const size_t mMax;
  ConstRef<IBufferFactory> mFactory;
  REF_COUNTED(BufferPoolFactory);
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
Result<IBufferRangeMutableCapacity> newBufferRange() noexcept {
  return makeRefResultNonNull<IBufferRangeMutableCapacity>(new (std::nothrow) Range());
}

Result<IBuffer> newOwnedBuffer(IAllocator* allocator, size_t size) noexcept {
  NON_NULL_PARAM_OR_RET(allocator);
  Ref<IMemory> memory;
  Ref<IBufferRangeMutableCapacity> range;
  UNWRAP_OR_FWD_RESULT(memory, allocator->allocate(size));
  UNWRAP_OR_FWD_RESULT(range, newBufferRange());
  range.get()->setCapacity(memory.get()->capacity());
  return makeRefResultNonNull<IBuffer>(new (std::nothrow) OwnedBuffer(memory.get(), range.get()));
}

Result<IBuffer> newBufferSlice(IBuffer* parent, size_t start, size_t end) noexcept {
  NON_NULL_PARAM_OR_RET(parent);
  GS_REQUIRE_OR_RET_RESULT_FMT(start <= end && end <= parent->range()->capacity(), "invalid slice [%zu, %zu) of capacity [%zu]", start, end,
                               parent->range()->capacity());
  Ref<IBufferRangeMutableCapacity> range;
  UNWRAP_OR_FWD_RESULT(range, newBufferRange());
  range.get()->setCapacity(end - start);
  const size_t pOff = parent->range()->offset(), pEnd = parent->range()->endOffset();
  // the slice's used range = the parent's used range seen through the window [start, end), relative to `start`
  // (reference BufferSlice.cpp:24-42): empty at 0 when the window lies behind the parent's data, empty at the window's
  // end when it lies in front of it
  const size_t len = end - start;
  const size_t sOff = pOff <= start ? 0 : pOff >= end ? len : pOff - start;
  const size_t sEnd = pEnd >= end ? len : pEnd <= start ? 0 : pEnd - start;
  FWD_IN_RESULT_IF_ERR(range.get()->setUsedRange(sOff < sEnd ? sOff : sEnd, sEnd));
  return makeRefResultNonNull<IBuffer>(new (std::nothrow) SliceBuffer(parent, start, range.get()));
}

IAllocator* newSysMemAllocator() noexcept { return new (std::nothrow) SysMemAllocator(); }
IBufferCopier* newSysMemCopier() noexcept { return new (std::nothrow) SysMemCopier(); }
IMemSet* newSysMemSet() noexcept { return new (std::nothrow) SysMemSet(); }
ICudaAllocatorFactory* newCudaAllocatorFactory() noexcept { return new (std::nothrow) CudaAllocatorFactory(); }
ICudaBufferCopierFactory* newCudaBufferCopierFactory() noexcept { return new (std::nothrow) CudaCopierFactory(); }
ICudaMemSetFactory* newCudaMemSetFactory() noexcept { return new (std::nothrow) CudaMemSetFactory(); }
IBufferRangeFactory* newBufferRangeFactory() noexcept { return new (std::nothrow) RangeFactory(); }
IBufferSliceFactory* newBufferSliceFactory() noexcept { return new (std::nothrow) SliceFactory(); }
IBufferUtil* newBufferUtil() noexcept { return new (std::nothrow) BufferUtil(); }
IBufferFactory* newBufferFactory(IAllocator* allocator) noexcept { return new (std::nothrow) BufferFactory(allocator); }
IResizableBufferFactory* newResizableBufferFactory(IAllocator* allocator, const IBufferCopier* copier) noexcept {
  return new (std::nothrow) ResizableFactory(allocator, copier);
}
IRelocatableResizableBufferFactory* newRelocatableBufferFactory(IAllocator* allocator, const IBufferCopier* copier) noexcept {
  return new (std::nothrow) RelocatableFactory(allocator, copier);
}
IBufferPool* newBufferPool(size_t maxBufferCount, size_t bufferSize, IBufferFactory* bufferFactory) noexcept {
  return new (std::nothrow) BufferPool(maxBufferCount, bufferSize, bufferFactory);
}
IBufferPoolFactory* newBufferPoolFactory(size_t maxBufferCount, IBufferFactory* bufferFactory) noexcept {
  return new (std::nothrow) BufferPoolFactory(maxBufferCount, bufferFactory);
}

}  // namespace gs
