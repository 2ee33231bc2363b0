/*
 * EventPipeline.h -- ADDITIVE: the one-deep event pipeline of the reference's output side as a public object.
 *
 * The reference keeps this helper private to its AAC sink (src/filters/Waiter.h, Waiter.cpp:34-50:
 * recordNextAndWaitPrevious() records an event behind the work just enqueued and blocks on the event recorded one call
 * earlier), which is what lets the device->host copy and the kernels of step i run while the host consumes the result
 * of step i-1.  Callers that read a device->host CudaMemcpyFilter into two alternating pinned buffers get the same
 * overlap with this object instead of a cudaStreamSynchronize per step (oracle/ref/ref_chain.cpp --pipeline 1).
 */
#ifndef GPUSDRPIPELINE_EVENTPIPELINE_H
#define GPUSDRPIPELINE_EVENTPIPELINE_H

#include <gpusdrpipeline/Factories.h>

class IEventPipeline : public virtual IRef {
 public:
  /* Record an event on the queue's stream behind everything enqueued so far, then wait (host-side) for the event the
   * PREVIOUS call recorded.  On return the work enqueued before the previous call is complete. */
  [[nodiscard]] virtual Status recordNextAndWaitPrevious() noexcept = 0;
  /* Wait for the event of the last call (end of stream). */
  [[nodiscard]] virtual Status waitLast() noexcept = 0;
  ABSTRACT_IREF(IEventPipeline);
};

GS_EXPORT [[nodiscard]] Result<IEventPipeline> gsCreateEventPipeline(ICudaCommandQueue* commandQueue) noexcept;

#endif  // GPUSDRPIPELINE_EVENTPIPELINE_H
