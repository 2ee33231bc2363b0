/*
 * queues.h -- command queues: a (CUDA device, CUDA stream) pair every node enqueues its work on.
 * Mirrors reference include/gpusdrpipeline/commandqueue/*.h (cited per interface).
 */
#ifndef GPUSDRPIPELINE_ABI_QUEUES_H
#define GPUSDRPIPELINE_ABI_QUEUES_H

#include <gpusdrpipeline/abi/buffers.h>

// commandqueue/ICommandQueue.h:23-26
class ICommandQueue : public virtual IRef {
 public:
  ABSTRACT_IREF(ICommandQueue);
};

// commandqueue/ICudaCommandQueue.h:23-29.  The stream is a BLOCKING stream (cudaStreamCreate): the reference's own
// tests synchronise through the legacy default stream (tests/CosineSourceTests.cpp:41-47).
class ICudaCommandQueue : public ICommandQueue {
 public:
  virtual int32_t cudaDevice() const noexcept = 0;
  virtual cudaStream_t cudaStream() const noexcept = 0;
  ABSTRACT_IREF(ICudaCommandQueue);
};

// commandqueue/ICudaCommandQueueFactory.h:11-16 (IRef is a NON-virtual base here, as in the reference)
class ICudaCommandQueueFactory : public IRef {
 public:
  virtual Result<ICudaCommandQueue> create(int32_t cudaDevice) noexcept = 0;
  ABSTRACT_IREF(ICudaCommandQueueFactory);
};

// commandqueue/ICommandQueueFactory.h:25-62 -- named queues created from JSON: {"queueType":"cuda","cudaDevice":N}
class ICommandQueueFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Status create(const char* queueId, const char* parameterJson) noexcept = 0;
  [[nodiscard]] virtual bool exists(const char* queueId) noexcept = 0;
  [[nodiscard]] virtual Result<ICudaCommandQueue> getCudaCommandQueue(const char* queueId) noexcept = 0;
  ABSTRACT_IREF(ICommandQueueFactory);
};

// commandqueue/IExecDevice.h:23-39 (IExecDevice / IExecDeviceList) is declared by the reference but implemented and used by
// nothing in its tree; it is not part of this boundary.

#endif  // GPUSDRPIPELINE_ABI_QUEUES_H
