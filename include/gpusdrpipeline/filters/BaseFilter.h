/*
 * BaseFilter.h -- BaseSink + BaseSource for out-of-tree Filters.  Mirrors reference
 * include/gpusdrpipeline/filters/BaseFilter.h:33-47 / src/filters/BaseFilter.cpp; exported (see BaseSink.h).
 * A derived filter implements getOutputDataSize / getOutputSizeAlignment / readOutput / preferredInputBufferSize, reads
 * its input through getPortInputBuffer(port) and finishes a step with consumeInputBytesAndMoveUsedToStart().
 */
#ifndef GPUSDRPIPELINE_FILTERS_BASEFILTER_H
#define GPUSDRPIPELINE_FILTERS_BASEFILTER_H

#include <gpusdrpipeline/filters/BaseSink.h>
#include <gpusdrpipeline/filters/BaseSource.h>

#include <cstdint>
#include <vector>

class GS_PUBLIC BaseFilter : public virtual Filter, public BaseSink, public BaseSource {
 public:
  BaseFilter() = delete;

 protected:
  BaseFilter(IRelocatableResizableBufferFactory* relocatableResizableBufferFactory, IBufferSliceFactory* slicedBufferFactory,
             size_t inputPortCount, std::vector<ImmutableRef<IBufferCopier>>&& outputPortBufferCopiers, IMemSet* memSet = nullptr) noexcept;
  ~BaseFilter() override = default;
};

#endif  // GPUSDRPIPELINE_FILTERS_BASEFILTER_H
