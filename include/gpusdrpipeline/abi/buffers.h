/*
 * buffers.h -- memory, range bookkeeping and buffer interfaces of the gpusdrpipeline boundary.
 *
 * One header for what the reference spreads over include/gpusdrpipeline/IMemory.h and 23 files under
 * include/gpusdrpipeline/buffers/ (each still exists as a forwarding header).  Class names, inheritance (including
 * which bases are virtual) and the order of the virtual functions are the reference's; the cited line is where each
 * interface is declared there.
 *
 * The buffer contract the hot path relies on (reference filters/Filter.h:43-68, src/filters/BaseSink.cpp:61-116):
 *   IBuffer      = base pointer + IBufferRange;  readPtr() = base + offset,  writePtr() = base + endOffset
 *   IBufferRange = [offset, endOffset) used bytes inside [0, capacity)
 */
#ifndef GPUSDRPIPELINE_ABI_BUFFERS_H
#define GPUSDRPIPELINE_ABI_BUFFERS_H

#include <gpusdrpipeline/abi/core.h>

class ICommandQueue;
class ICudaCommandQueue;

// IMemory.h:24-43 -- an owned allocation
class IMemory : public virtual IRef {
 public:
  [[nodiscard]] virtual uint8_t* data() noexcept = 0;
  [[nodiscard]] virtual const uint8_t* data() const noexcept = 0;
  [[nodiscard]] virtual size_t capacity() const noexcept = 0;

  template <typename T = uint8_t>
  [[nodiscard]] T* as() noexcept {
    return reinterpret_cast<T*>(data());
  }
  template <typename T = uint8_t>
  [[nodiscard]] const T* as() const noexcept {
    return reinterpret_cast<const T*>(data());
  }
  ABSTRACT_IREF(IMemory);
};

// buffers/IAllocator.h:30-41
class IAllocator : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IMemory> allocate(size_t size) noexcept = 0;
  ABSTRACT_IREF(IAllocator);
};

// buffers/IAllocatorFactory.h:24-29
class IAllocatorFactory : public virtual IRef {
 public:
  virtual Result<IAllocator> create(ICommandQueue* forCommandQueue) = 0;
  ABSTRACT_IREF(IAllocatorFactory);
};

// buffers/IBufferCopier.h:25-30
class IBufferCopier : public virtual IRef {
 public:
  [[nodiscard]] virtual Status copy(void* dst, const void* src, size_t length) const noexcept = 0;
  ABSTRACT_IREF(IBufferCopier);
};

// buffers/IMemSet.h:25-30
class IMemSet : public virtual IRef {
 public:
  [[nodiscard]] virtual Status memSet(void* data, uint8_t value, size_t byteCount) noexcept = 0;
  ABSTRACT_IREF(IMemSet);
};

// buffers/IBufferRange.h:29-84
class IBufferRange : public virtual IRef {
 public:
  [[nodiscard]] virtual size_t capacity() const noexcept = 0;   // bytes reachable from IBuffer::base()
  [[nodiscard]] virtual size_t offset() const noexcept = 0;     // first used byte
  [[nodiscard]] virtual size_t endOffset() const noexcept = 0;  // one past the last used byte
  [[nodiscard]] virtual Status setUsedRange(size_t offset, size_t endOffset) noexcept = 0;
  [[nodiscard]] virtual size_t used() const noexcept { return endOffset() - offset(); }
  [[nodiscard]] virtual size_t remaining() const noexcept { return capacity() - endOffset(); }
  [[nodiscard]] virtual bool hasRemaining() const noexcept { return remaining() > 0; }

  void clearRange() noexcept { (void)setUsedRange(0, 0); }
  [[nodiscard]] Status increaseOffset(size_t increaseBy) {
    const size_t start = offset() + increaseBy, end = endOffset();
    if (start > end) {
      gsloge("New start offset [%zu] exceeds the end offset [%zu]", start, end);
      return Status_InvalidArgument;
    }
    return setUsedRange(start, end);
  }
  [[nodiscard]] Status increaseEndOffset(size_t increaseBy) {
    const size_t end = endOffset() + increaseBy, cap = capacity();
    if (end > cap) {
      gsloge("New end offset [%zu] exceeds the capacity [%zu]", end, cap);
      return Status_InvalidArgument;
    }
    return setUsedRange(offset(), end);
  }
  ABSTRACT_IREF(IBufferRange);
};

// buffers/IBufferRangeMutableCapacity.h:22-27
class IBufferRangeMutableCapacity : public IBufferRange {
 public:
  virtual void setCapacity(size_t capacity) noexcept = 0;
  ABSTRACT_IREF(IBufferRangeMutableCapacity);
};

// buffers/IBufferRangeFactory.h:23-38
class IBufferRangeFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IBufferRangeMutableCapacity> createBufferRange() const noexcept = 0;
  [[nodiscard]] Result<IBufferRangeMutableCapacity> createBufferRangeWithCapacity(size_t capacity) const {
    IBufferRangeMutableCapacity* range;
    UNWRAP_OR_FWD_RESULT(range, createBufferRange());
    range->setCapacity(capacity);
    return makeRefResultNonNull(range);
  }

 protected:
  ABSTRACT_IREF(IBufferRangeFactory);
};

// buffers/IBuffer.h:26-48
class IBuffer : public virtual IRef {
 public:
  [[nodiscard]] virtual uint8_t* base() noexcept = 0;
  [[nodiscard]] virtual const uint8_t* base() const noexcept = 0;
  [[nodiscard]] virtual IBufferRange* range() noexcept = 0;
  [[nodiscard]] virtual const IBufferRange* range() const noexcept = 0;

  template <class T = uint8_t>
  [[nodiscard]] const T* readPtr() const noexcept {
    return reinterpret_cast<const T*>(base() + range()->offset());
  }
  template <class T = uint8_t>
  [[nodiscard]] T* writePtr() noexcept {
    return reinterpret_cast<T*>(base() + range()->endOffset());
  }
  ABSTRACT_IREF(IBuffer);
};

// buffers/IBufferFactory.h:25-30
class IBufferFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IBuffer> createBuffer(size_t size) noexcept = 0;
  ABSTRACT_IREF(IBufferFactory);
};

// buffers/IBufferSliceFactory.h:23-57 -- a view [sliceStartOffset, sliceEndOffset) of another buffer with its own range
class IBufferSliceFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IBuffer> slice(IBuffer* bufferToSlice, size_t sliceStartOffset, size_t sliceEndOffset) noexcept = 0;
  [[nodiscard]] Result<IBuffer> sliceRemaining(IBuffer* bufferToSlice) {
    return slice(bufferToSlice, bufferToSlice->range()->endOffset(), bufferToSlice->range()->capacity());
  }
  ABSTRACT_IREF(IBufferSliceFactory);
};

// buffers/IBufferUtil.h:23-44
class IBufferUtil : public virtual IRef {
 public:
  [[nodiscard]] virtual Status appendToBuffer(IBuffer* buffer, const void* src, size_t count, const IBufferCopier* bufferCopier) const noexcept = 0;
  [[nodiscard]] virtual Status readFromBuffer(void* dst, IBuffer* buffer, size_t count, const IBufferCopier* bufferCopier) const noexcept = 0;
  [[nodiscard]] virtual Status moveFromBuffer(IBuffer* dst, IBuffer* src, size_t count, const IBufferCopier* bufferCopier) const noexcept = 0;
  ABSTRACT_IREF(IBufferUtil);
};

// buffers/IBufferPool.h:24-48 -- fixed-size buffers that return to the pool when the last reference goes away
class IBufferPool : public virtual IRef {
 public:
  [[nodiscard]] virtual size_t getBufferSize() const noexcept = 0;
  [[nodiscard]] virtual Result<IBuffer> getBuffer() noexcept = 0;     // blocks until a buffer is free
  [[nodiscard]] virtual Result<IBuffer> tryGetBuffer() noexcept = 0;  // value == nullptr when none is free
  ABSTRACT_IREF(IBufferPool);
};

// buffers/IBufferPoolFactory.h:22-27
class IBufferPoolFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IBufferPool> createBufferPool(size_t bufferSize) noexcept = 0;
  ABSTRACT_IREF(IBufferPoolFactory);
};

// buffers/IRelocatable.h:22-27, IResizable.h:24-29
class IRelocatable : public virtual IRef {
 public:
  [[nodiscard]] virtual Status relocate(size_t dstOffset, size_t srcOffset, size_t length) noexcept = 0;
  ABSTRACT_IREF(IRelocatable);
};
class IResizable : public virtual IRef {
 public:
  [[nodiscard]] virtual Status resize(size_t newSize) noexcept = 0;
  ABSTRACT_IREF(IResizable);
};

// buffers/IResizableBuffer.h:23-34
class IResizableBuffer : public IBuffer, public IResizable {
 public:
  [[nodiscard]] Status ensureMinSize(size_t minSize) noexcept { return range()->capacity() < minSize ? resize(minSize) : Status_Success; }
  ABSTRACT_IREF(IResizableBuffer);
};

// buffers/IResizableBufferFactory.h:24-29
class IResizableBufferFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IResizableBuffer> createResizableBuffer(size_t size) noexcept = 0;
  ABSTRACT_IREF(IResizableBufferFactory);
};

// buffers/IRelocatableResizableBuffer.h:23-28
class IRelocatableResizableBuffer : public IRelocatable, public IResizableBuffer {
 public:
  [[nodiscard]] Status relocateUsedToStart() noexcept { return relocate(0, range()->offset(), range()->used()); }
  ABSTRACT_IREF(IRelocatableResizableBuffer);
};

// buffers/IRelocatableResizableBufferFactory.h:22-27
class IRelocatableResizableBufferFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IRelocatableResizableBuffer> createRelocatableBuffer(size_t size) const noexcept = 0;
  ABSTRACT_IREF(IRelocatableResizableBufferFactory);
};

// buffers/IRelocatableCudaBufferFactory.h:24-33
class IRelocatableCudaBufferFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IRelocatableResizableBuffer> createCudaBuffer(size_t minSize, ICudaCommandQueue* commandQueue,
                                                                             size_t alignment, bool useHostMemory) noexcept = 0;
  ABSTRACT_IREF(IRelocatableCudaBufferFactory);
};

// buffers/ICudaAllocatorFactory.h:25-33 -- device memory (stream-ordered) or pinned host memory
class ICudaAllocatorFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IAllocator> createCudaAllocator(ICudaCommandQueue* commandQueue, size_t alignment, bool useHostMemory) noexcept = 0;
  ABSTRACT_IREF(ICudaAllocatorFactory);
};

// buffers/ICudaBufferCopierFactory.h:27-34 -- cudaMemcpyAsync of the given kind on the queue's stream
class ICudaBufferCopierFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IBufferCopier> createBufferCopier(ICudaCommandQueue* commandQueue, cudaMemcpyKind memcpyKind) noexcept = 0;
  ABSTRACT_IREF(ICudaBufferCopierFactory);
};

// buffers/ICudaMemSetFactory.h:30-35
class ICudaMemSetFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IMemSet> create(ICudaCommandQueue* commandQueue) noexcept = 0;
  ABSTRACT_IREF(ICudaMemSetFactory);
};

#endif  // GPUSDRPIPELINE_ABI_BUFFERS_H
