/*
 * BaseSink.h -- helper base class for out-of-tree Sinks: per-port input accumulation behind requestBuffer()/commitBuffer().
 *
 * Mirrors reference include/gpusdrpipeline/filters/BaseSink.h:30-76 (same class name, bases, member order and protected
 * interface, so a filter written against the reference compiles unchanged); the behaviour is that of
 * src/filters/BaseSink.cpp:61-170.  The nodes of THIS library do not derive from it -- they keep their input in a
 * PortInput (lazy compaction, cuda_sdr_b200/host/port_input.h); this class is for user-written nodes and keeps the
 * reference's contract: consumeInputBytesAndMoveUsedToStart() relocates what is left to the start of the buffer.
 *
 * Unlike the reference build (hidden visibility, class not exported: src/CMakeLists.txt:263-264) the class is exported
 * from libgpusdrpipeline.so, so deriving from it links.
 */
#ifndef GPUSDRPIPELINE_FILTERS_BASESINK_H
#define GPUSDRPIPELINE_FILTERS_BASESINK_H

#include <gpusdrpipeline/Factories.h>

#include <vector>

class GS_PUBLIC BaseSink : public virtual Sink {
 public:
  struct InputPort {
    ConstRef<IRelocatableResizableBuffer> inputBuffer;
    bool bufferCheckedOut;
  };

  BaseSink() = delete;
  [[nodiscard]] Result<IBuffer> requestBuffer(size_t port, size_t numBytes) noexcept override;
  [[nodiscard]] Status commitBuffer(size_t port, size_t byteCount) noexcept override;

 protected:
  BaseSink(IRelocatableResizableBufferFactory* relocatableResizableBufferFactory, IBufferSliceFactory* slicedBufferFactory,
           size_t inputPortCount, IMemSet* memSet = nullptr);
  ~BaseSink() override = default;

  [[nodiscard]] Result<IBuffer> getPortInputBuffer(size_t port) noexcept;
  [[nodiscard]] Result<const IBuffer> getPortInputBuffer(size_t port) const noexcept;
  [[nodiscard]] bool inputPortsInitialized() const noexcept;
  /* advances the port buffer's offset by numBytes, then moves the unconsumed bytes to the start of the buffer */
  [[nodiscard]] Status consumeInputBytesAndMoveUsedToStart(size_t port, size_t numBytes) noexcept;

 private:
  const size_t mInputPortCount;
  ConstRef<IBufferSliceFactory> mSlicedBufferFactory;
  std::vector<InputPort> mInputPorts;
  ConstRef<IMemSet> mMemSet;
  ConstRef<IRelocatableResizableBufferFactory> mRelocatableResizableBufferFactory;

  [[nodiscard]] Status ensureInputPortsInit() noexcept;
};

#endif  // GPUSDRPIPELINE_FILTERS_BASESINK_H
