// Internal declarations shared by the translation units of libgpusdrpipeline.so (the C++ host framework that mirrors
// the reference's getFactoriesSingleton()/IFactories boundary on top of the sm_100a kernels in libb200sdr.so).
#pragma once

#include <gpusdrpipeline/Factories.h>

#include <string>
#include <vector>

namespace gs {

inline size_t roundUp(size_t value, size_t multiple) { return multiple == 0 ? value : (value + multiple - 1) / multiple * multiple; }

// ---- buffers.cpp --------------------------------------------------------------------------------------------
Result<IBufferRangeMutableCapacity> newBufferRange() noexcept;
Result<IBuffer> newOwnedBuffer(IAllocator* allocator, size_t size) noexcept;
Result<IBuffer> newBufferSlice(IBuffer* parent, size_t start, size_t end) noexcept;
IAllocator* newSysMemAllocator() noexcept;
IBufferCopier* newSysMemCopier() noexcept;
IMemSet* newSysMemSet() noexcept;
ICudaAllocatorFactory* newCudaAllocatorFactory() noexcept;
ICudaBufferCopierFactory* newCudaBufferCopierFactory() noexcept;
ICudaMemSetFactory* newCudaMemSetFactory() noexcept;
IBufferRangeFactory* newBufferRangeFactory() noexcept;
IBufferSliceFactory* newBufferSliceFactory() noexcept;
IBufferUtil* newBufferUtil() noexcept;
IBufferFactory* newBufferFactory(IAllocator* allocator) noexcept;
IResizableBufferFactory* newResizableBufferFactory(IAllocator* allocator, const IBufferCopier* copier) noexcept;
IRelocatableResizableBufferFactory* newRelocatableBufferFactory(IAllocator* allocator, const IBufferCopier* copier) noexcept;
IBufferPool* newBufferPool(size_t maxBufferCount, size_t bufferSize, IBufferFactory* bufferFactory) noexcept;
IBufferPoolFactory* newBufferPoolFactory(size_t maxBufferCount, IBufferFactory* bufferFactory) noexcept;

// ---- queues.cpp ---------------------------------------------------------------------------------------------
ICudaCommandQueueFactory* newCudaCommandQueueFactory() noexcept;
ICommandQueueFactory* newCommandQueueFactory(IFactories* factories) noexcept;

// ---- nodes.cpp / fused_node.cpp -----------------------------------------------------------------------------
ICudaMemcpyFilterFactory* newCudaMemcpyFilterFactory(IFactories* f) noexcept;
IAacFileWriterFactory* newAacFileWriterFactory() noexcept;
IAddConstFactory* newAddConstFactory(IFactories* f) noexcept;
IAddConstToVectorLengthFactory* newAddConstToVectorLengthFactory(IFactories* f) noexcept;
ICosineSourceFactory* newCosineSourceFactory(IFactories* f) noexcept;
IFileReaderFactory* newFileReaderFactory() noexcept;
IFirFactory* newFirFactory(IFactories* f) noexcept;
IHackrfSourceFactory* newHackrfSourceFactory() noexcept;
ICudaFilterFactory* newInt8ToFloatFactory(IFactories* f) noexcept;
ICudaFilterFactory* newMagnitudeFactory(IFactories* f) noexcept;
ICudaFilterFactory* newMultiplyFactory(IFactories* f) noexcept;
IQuadDemodFactory* newQuadDemodFactory(IFactories* f) noexcept;
IPortRemappingSinkFactory* newPortRemappingSinkFactory() noexcept;
IPortRemappingSourceFactory* newPortRemappingSourceFactory() noexcept;
IReadByteCountMonitorFactory* newReadByteCountMonitorFactory() noexcept;
IRfToPcmAudioFactory* newRfToPcmAudioFactory(IFactories* f) noexcept;

// ---- drivers.cpp --------------------------------------------------------------------------------------------
ISteppingDriverFactory* newSteppingDriverFactory() noexcept;
IFilterDriverFactory* newFilterDriverFactory(IFactories* f) noexcept;
IDriverToDiagramFactory* newDriverToDotFactory() noexcept;

}  // namespace gs
