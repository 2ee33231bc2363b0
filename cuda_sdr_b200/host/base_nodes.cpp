// BaseSink / BaseSource / BaseFilter -- the helper base classes the reference publishes for user-written nodes
// (include/gpusdrpipeline/filters/Base{Sink,Source,Filter}.h; behaviour: reference src/filters/BaseSink.cpp:61-170,
// BaseSource.cpp, BaseFilter.cpp).  The library's own nodes do not use them (host/port_input.h); they are exported so
// that an out-of-tree filter deriving from BaseFilter compiles and links against libgpusdrpipeline.so.
#include <gpusdrpipeline/filters/BaseFilter.h>

#include "internal.h"

namespace {
constexpr size_t kInitialPortBytes = 8192;  // BaseSink.cpp:50
}

BaseSink::BaseSink(IRelocatableResizableBufferFactory* relocatableResizableBufferFactory, IBufferSliceFactory* slicedBufferFactory,
                   size_t inputPortCount, IMemSet* memSet)
    : mInputPortCount(inputPortCount),
      mSlicedBufferFactory(slicedBufferFactory),
      mMemSet(memSet),
      mRelocatableResizableBufferFactory(relocatableResizableBufferFactory) {}

// port buffers are created on first use: a constructor cannot report an allocation failure
Status BaseSink::ensureInputPortsInit() noexcept {
  if (mInputPorts.size() >= mInputPortCount) return Status_Success;
  GS_REQUIRE_OR_RET_STATUS(mRelocatableResizableBufferFactory != nullptr && mSlicedBufferFactory != nullptr,
                           "BaseSink needs a buffer factory and a slice factory");
  try {
    mInputPorts.reserve(mInputPortCount);
    while (mInputPorts.size() < mInputPortCount) {
      Ref<IRelocatableResizableBuffer> buffer;
      UNWRAP_OR_FWD_STATUS(buffer, mRelocatableResizableBufferFactory->createRelocatableBuffer(kInitialPortBytes));
      mInputPorts.push_back(InputPort {buffer.get(), false});
    }
    return Status_Success;
  }
  IF_CATCH_RETURN_STATUS
}

Result<IBuffer> BaseSink::requestBuffer(size_t port, size_t numBytes) noexcept {
  FWD_IN_RESULT_IF_ERR(ensureInputPortsInit());
  GS_REQUIRE_OR_RET_RESULT_FMT(port < mInputPorts.size(), "Cannot request buffer. Input port [%zu] is out of range.", port);
  InputPort& in = mInputPorts[port];
  GS_REQUIRE_OR_RET_RESULT(!in.bufferCheckedOut, "Cannot request buffer - it is already checked out");
  IRelocatableResizableBuffer* buffer = in.inputBuffer.get();
  if (buffer->range()->remaining() < numBytes) FWD_IN_RESULT_IF_ERR(buffer->resize(buffer->range()->endOffset() + numBytes));
  IBuffer* view;
  UNWRAP_OR_FWD_RESULT(view, mSlicedBufferFactory->sliceRemaining(buffer));
  in.bufferCheckedOut = true;
  return makeRefResultNonNull(view);
}

Status BaseSink::commitBuffer(size_t port, size_t byteCount) noexcept {
  FWD_IF_ERR(ensureInputPortsInit());
  GS_REQUIRE_OR_RET_STATUS_FMT(port < mInputPorts.size(), "Cannot commit buffer. Input port [%zu] is out of range", port);
  InputPort& in = mInputPorts[port];
  GS_REQUIRE_OR_RET_STATUS(in.bufferCheckedOut, "Cannot commit buffer - it was not checked out");
  GS_REQUIRE_OR_RET_STATUS(byteCount <= in.inputBuffer->range()->remaining(),
                           "Cannot commit buffer - the committed number of bytes exceeds its capacity");
  FWD_IF_ERR(in.inputBuffer->range()->increaseEndOffset(byteCount));
  in.bufferCheckedOut = false;
  return Status_Success;
}

Result<IBuffer> BaseSink::getPortInputBuffer(size_t port) noexcept {
  FWD_IN_RESULT_IF_ERR(ensureInputPortsInit());
  GS_REQUIRE_OR_RET_RESULT_FMT(port < mInputPorts.size(), "Cannot get input buffer - Input port [%zu] is out of range", port);
  GS_REQUIRE_OR_RET_RESULT(!mInputPorts[port].bufferCheckedOut, "Cannot get input buffer - buffer is checked out");
  return makeRefResultNonNull<IBuffer>(mInputPorts[port].inputBuffer.get());
}

Result<const IBuffer> BaseSink::getPortInputBuffer(size_t port) const noexcept {
  GS_REQUIRE_OR_RET_RESULT_FMT(port < mInputPortCount, "Cannot get input buffer - Input port [%zu] is out of range", port);
  GS_REQUIRE_OR_RET_RESULT(port < mInputPorts.size(), "Cannot get input buffer - input buffer has not been created. Call getPortInputBuffer() first.");
  GS_REQUIRE_OR_RET_RESULT(!mInputPorts[port].bufferCheckedOut, "Cannot get input buffer - buffer is checked out");
  return makeRefResultNonNull<const IBuffer>(mInputPorts[port].inputBuffer.get());
}

bool BaseSink::inputPortsInitialized() const noexcept { return mInputPorts.size() == mInputPortCount; }

Status BaseSink::consumeInputBytesAndMoveUsedToStart(size_t port, size_t numBytes) noexcept {
  FWD_IF_ERR(ensureInputPortsInit());
  GS_REQUIRE_OR_RET_STATUS_FMT(port < mInputPorts.size(), "Cannot consume input - Input port [%zu] is out of range", port);
  IRelocatableResizableBuffer* buffer = mInputPorts[port].inputBuffer.get();
  if (numBytes == 0 && buffer->range()->offset() == 0) return Status_Success;
  FWD_IF_ERR(buffer->range()->increaseOffset(numBytes));
  return buffer->relocateUsedToStart();
}

BaseSource::BaseSource(std::vector<ImmutableRef<IBufferCopier>>&& outputPortBufferCopiers) noexcept
    : mOutputPortBufferCopiers(std::move(outputPortBufferCopiers)) {}

IBufferCopier* BaseSource::getOutputCopier(size_t port) noexcept {
  return port < mOutputPortBufferCopiers.size() ? mOutputPortBufferCopiers[port].get() : nullptr;
}

BaseFilter::BaseFilter(IRelocatableResizableBufferFactory* relocatableResizableBufferFactory, IBufferSliceFactory* slicedBufferFactory,
                       size_t inputPortCount, std::vector<ImmutableRef<IBufferCopier>>&& outputPortBufferCopiers, IMemSet* memSet) noexcept
    : BaseSink(relocatableResizableBufferFactory, slicedBufferFactory, inputPortCount, memSet), BaseSource(std::move(outputPortBufferCopiers)) {}
