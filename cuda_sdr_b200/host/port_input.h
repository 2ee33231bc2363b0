// PortInput -- the stream-state carrier of one input port: accumulates committed bytes, hands out the space behind them
// for the next requestBuffer(), and lets the node consume from the front.
//
// Replaces the reference's per-port RelocatableResizableBuffer (two allocations; after EVERY readOutput the unconsumed
// remainder is copied device-to-device into the twin and the two are swapped: BaseSink.cpp:150-170,
// RelocatableResizableBuffer.cpp:79-103).  Here consuming is pointer arithmetic; the remainder (FIR history, the FM
// discriminator's one sample, an unconsumed tail) is re-read in place by the next launch and is moved to the front of
// the allocation only when the next request no longer fits behind it.
//
// Host (pinned) ports -- the input side of a host->device copy node -- alternate between two allocations, each guarded
// by an event recorded after the copy that drained it, so the host may fill one while the other is still being copied.
#pragma once

#include <gpusdrpipeline/Factories.h>

namespace gs {

class PortInput {
 public:
  PortInput(IAllocator* allocator, const IBufferCopier* copier, ICudaCommandQueue* queue, bool host) noexcept;
  ~PortInput() noexcept;
  PortInput(const PortInput&) = delete;
  PortInput& operator=(const PortInput&) = delete;

  [[nodiscard]] Result<IBuffer> request(size_t bytes) noexcept;  // Sink::requestBuffer
  [[nodiscard]] Status commit(size_t bytes) noexcept;            // Sink::commitBuffer
  [[nodiscard]] const uint8_t* data() const noexcept { return mMemory != nullptr ? mMemory.get()->data() + mOffset : nullptr; }
  [[nodiscard]] size_t used() const noexcept { return mEnd - mOffset; }
  void consume(size_t bytes) noexcept;
  // host ports: call after enqueuing the copy that reads the consumed bytes
  [[nodiscard]] Status fenceDrained() noexcept;
  [[nodiscard]] bool isHost() const noexcept { return mHost; }

 private:
  [[nodiscard]] Status makeRoom(size_t bytes) noexcept;

  ConstRef<IAllocator> mAllocator;
  ConstRef<const IBufferCopier> mCopier;
  ConstRef<ICudaCommandQueue> mQueue;
  const bool mHost;
  Ref<IMemory> mMemory;
  size_t mCapacity = 0, mOffset = 0, mEnd = 0;
  bool mCheckedOut = false;
  // host double buffering
  Ref<IMemory> mOther;
  cudaEvent_t mFence = nullptr, mOtherFence = nullptr;
  bool mFencePending = false, mOtherFencePending = false;
};

}  // namespace gs
