/*
 * BaseSource.h -- helper base class for out-of-tree Sources: one output copier per port (fan-out of a port to several
 * sinks).  Mirrors reference include/gpusdrpipeline/filters/BaseSource.h:24-35 / src/filters/BaseSource.cpp; exported
 * from libgpusdrpipeline.so (see BaseSink.h).
 */
#ifndef GPUSDRPIPELINE_FILTERS_BASESOURCE_H
#define GPUSDRPIPELINE_FILTERS_BASESOURCE_H

#include <gpusdrpipeline/Factories.h>

#include <vector>

class GS_PUBLIC BaseSource : public virtual Source {
 public:
  explicit BaseSource(std::vector<ImmutableRef<IBufferCopier>>&& outputPortBufferCopiers) noexcept;
  IBufferCopier* getOutputCopier(size_t port) noexcept override;

 protected:
  ~BaseSource() override = default;

 private:
  const std::vector<ImmutableRef<IBufferCopier>> mOutputPortBufferCopiers;
};

#endif  // GPUSDRPIPELINE_FILTERS_BASESOURCE_H
