/*
 * nodes.h -- the graph-node contract (Node / Sink / Source / Filter) and the typed node factories.
 *
 * Mirrors reference include/gpusdrpipeline/filters/{Filter,FilterFactories,IHackrfSource,IPortRemappingSink,
 * IPortRemappingSource,IReadByteCountMonitor}.h -- same class names, same (virtual) inheritance, same order of the
 * virtual functions, so objects cross the library boundary in either direction.
 *
 * The buffer contract (reference filters/Filter.h:43-132, src/filters/BaseSink.cpp:61-170):
 *   Sink::requestBuffer(port, n)  -> an IBuffer with >= n writable bytes at writePtr() (device memory for GPU nodes,
 *                                    pinned host memory for a host->device copy node); one outstanding request per port
 *   Sink::commitBuffer(port, n)   -> the first n bytes written are now part of the port's input; 0 cancels
 *   Source::getOutputDataSize     -> bytes the next readOutput() would produce given unlimited room
 *   Source::readOutput(bufs, n)   -> appends at bufs[p]->writePtr(), advances its range's endOffset; produces less when
 *                                    room is short and then neither loses nor skips input (tests/FirTests.cpp:96-221)
 * All GPU work is enqueued on the node's ICudaCommandQueue stream; nothing synchronises.
 */
#ifndef GPUSDRPIPELINE_ABI_NODES_H
#define GPUSDRPIPELINE_ABI_NODES_H

#include <gpusdrpipeline/abi/queues.h>

class Sink;
class Source;
class Filter;
class IDriver;

// filters/Filter.h:30-40
class Node : public virtual IRef {
 public:
  virtual Sink* asSink() noexcept { return nullptr; }
  virtual Source* asSource() noexcept { return nullptr; }
  virtual Filter* asFilter() noexcept { return nullptr; }
  virtual IDriver* asDriver() noexcept { return nullptr; }
  virtual void updateParameters(const char* jsonParameters) noexcept {};
  ABSTRACT_IREF(Node);
};

// filters/Filter.h:41-73
class Sink : public virtual Node {
 public:
  [[nodiscard]] virtual Result<IBuffer> requestBuffer(size_t port, size_t byteCount) noexcept = 0;
  [[nodiscard]] virtual Status commitBuffer(size_t port, size_t byteCount) noexcept = 0;
  [[nodiscard]] virtual size_t preferredInputBufferSize(size_t port) noexcept = 0;
  [[nodiscard]] Sink* asSink() noexcept override { return this; }
  ABSTRACT_IREF(Sink);
};

// filters/Filter.h:75-132
class Source : public virtual Node {
 public:
  [[nodiscard]] virtual size_t getOutputDataSize(size_t port) noexcept = 0;
  [[nodiscard]] virtual size_t getOutputSizeAlignment(size_t port) noexcept = 0;
  [[nodiscard]] virtual IBufferCopier* getOutputCopier(size_t port) noexcept = 0;
  // getOutputDataSize() rounded up to the alignment (saturating)
  [[nodiscard]] virtual size_t getAlignedOutputDataSize(size_t port) noexcept {
    const size_t alignment = getOutputSizeAlignment(port);
    const size_t size = getOutputDataSize(port);
    if (alignment == 0) return size;
    if (size > SIZE_MAX - alignment + 1) return size / alignment * alignment;
    return (size + alignment - 1) / alignment * alignment;
  }
  [[nodiscard]] virtual Status readOutput(IBuffer** portOutputBuffers, size_t numPorts) noexcept = 0;
  [[nodiscard]] Source* asSource() noexcept override { return this; }
  ABSTRACT_IREF(Source);
};

// filters/Filter.h:134-138
class Filter : public virtual Sink, public virtual Source {
  ABSTRACT_IREF(Filter);
  Filter* asFilter() noexcept override { return this; }
};

// filters/IHackrfSource.h:27-41 (hardware source; this library's factory for it answers Status_NotFound)
class IHackrfSource : public virtual Source {
 public:
  [[nodiscard]] virtual int32_t getDeviceCount() const noexcept = 0;
  [[nodiscard]] virtual size_t getDeviceSerialNumber(int32_t deviceIndex, char* buffer, size_t bufferSize) const noexcept = 0;
  [[nodiscard]] virtual Status selectDeviceByIndex(int32_t deviceIndex) noexcept = 0;
  [[nodiscard]] virtual Status selectDeviceBySerialNumber(const char* serialNumber) noexcept = 0;
  [[nodiscard]] virtual Status releaseDevice() noexcept = 0;
  [[nodiscard]] virtual Status start() noexcept = 0;
  [[nodiscard]] virtual Status stop() noexcept = 0;
  ABSTRACT_IREF(IHackrfSource);
};

// filters/IPortRemappingSink.h:22-27, IPortRemappingSource.h:22-27
class IPortRemappingSink : public virtual Sink {
 public:
  virtual void addPortMapping(size_t outerPort, Sink* innerSink, size_t innerSinkPort) noexcept = 0;
  ABSTRACT_IREF(IPortRemappingSink);
};
class IPortRemappingSource : public virtual Source {
 public:
  virtual void addPortMapping(size_t outerPort, Source* innerSource, size_t innerSourcePort) noexcept = 0;
  ABSTRACT_IREF(IPortRemappingSource);
};

// filters/IReadByteCountMonitor.h:24-29
class IReadByteCountMonitor : public Filter {
 public:
  [[nodiscard]] virtual size_t getByteCountRead(size_t port) noexcept = 0;
  ABSTRACT_IREF(IReadByteCountMonitor);
};

/* ---- factories (filters/FilterFactories.h:30-182) --------------------------------------------------------- */
class INodeFactory : public virtual IRef {
 public:
  virtual Result<Node> create(const char* jsonParameters) noexcept = 0;
  ABSTRACT_IREF(INodeFactory);
};

// name -> factory registry (FilterFactories.h:36-43).  The reference DEFINES registerDefaultFilterFactories
// (src/filters/FilterFactories.cpp:132) while declaring registerDefaultNodeFactories; this library exports both.
GS_EXPORT [[nodiscard]] Result<Node> createNode(const char* name, const char* jsonParameters) noexcept;
GS_EXPORT [[nodiscard]] Result<Filter> createFilter(const char* name, const char* jsonParameters) noexcept;
GS_EXPORT [[nodiscard]] Result<Source> createSource(const char* name, const char* jsonParameters) noexcept;
GS_EXPORT [[nodiscard]] Result<Sink> createSink(const char* name, const char* jsonParameters) noexcept;
GS_EXPORT [[nodiscard]] bool hasNodeFactory(const char* name) noexcept;
GS_EXPORT [[nodiscard]] Status registerNodeFactory(const char* name, INodeFactory* filterFactory) noexcept;
GS_EXPORT [[nodiscard]] Status registerDefaultNodeFactories() noexcept;
GS_EXPORT [[nodiscard]] Status registerDefaultFilterFactories() noexcept;

class ICudaMemcpyFilterFactory : public INodeFactory {
 public:
  [[nodiscard]] virtual Result<Filter> createCudaMemcpy(cudaMemcpyKind memcpyKind, ICudaCommandQueue* commandQueue) noexcept = 0;
  ABSTRACT_IREF(ICudaMemcpyFilterFactory);
};
class IAacFileWriterFactory : public INodeFactory {
 public:
  [[nodiscard]] virtual Result<Sink> createAacFileWriter(const char* outputFileName, int32_t sampleRate, int32_t bitRate,
                                                         ICudaCommandQueue* commandQueue) noexcept = 0;
  ABSTRACT_IREF(IAacFileWriterFactory);
};
class IAddConstFactory : public INodeFactory {
 public:
  [[nodiscard]] virtual Result<Filter> createAddConst(float addValueToAmplitude, ICudaCommandQueue* commandQueue) noexcept = 0;
  ABSTRACT_IREF(IAddConstFactory);
};
class IAddConstToVectorLengthFactory : public INodeFactory {
 public:
  [[nodiscard]] virtual Result<Filter> createAddConstToVectorLength(float addValueToMagnitude, ICudaCommandQueue* commandQueue) noexcept = 0;
  ABSTRACT_IREF(IAddConstToVectorLengthFactory);
};
class ICosineSourceFactory : public INodeFactory {
 public:
  [[nodiscard]] virtual Result<Source> createCosineSource(SampleType sampleType, float sampleRate, float frequency,
                                                          ICudaCommandQueue* commandQueue) noexcept = 0;
  ABSTRACT_IREF(ICosineSourceFactory);
};
class IFileReaderFactory : public INodeFactory {
 public:
  [[nodiscard]] virtual Result<Source> createFileReader(const char* fileName) noexcept = 0;
  ABSTRACT_IREF(IFileReaderFactory);
};
class IFirFactory : public INodeFactory {
 public:
  // taps are used as given, in correlation order: out[k] = sum_j taps[j] * in[k*decimation + j]
  [[nodiscard]] virtual Result<Filter> createFir(SampleType tapType, SampleType elementType, size_t decimation, const float* taps,
                                                 size_t tapCount, ICudaCommandQueue* commandQueue) noexcept = 0;
  ABSTRACT_IREF(IFirFactory);
};
class IHackrfSourceFactory : public INodeFactory {
 public:
  [[nodiscard]] virtual Result<IHackrfSource> createHackrfSource(int32_t deviceIndex, uint64_t centerFrequency, double sampleRate,
                                                                 size_t maxBufferCountBeforeDropping) noexcept = 0;
  ABSTRACT_IREF(IHackrfSourceFactory);
};
class ICudaFilterFactory : public INodeFactory {
 public:
  [[nodiscard]] virtual Result<Filter> createFilter(ICudaCommandQueue* commandQueue) noexcept = 0;
  ABSTRACT_IREF(ICudaFilterFactory);
};
class IPortRemappingSinkFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IPortRemappingSink> create() noexcept = 0;
  ABSTRACT_IREF(IPortRemappingSinkFactory);
};
class IPortRemappingSourceFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IPortRemappingSource> create() noexcept = 0;
  ABSTRACT_IREF(IPortRemappingSourceFactory);
};
class IQuadDemodFactory : public INodeFactory {
 public:
  // fskDeviation is only used for FM; gain = rfSampleRate / (2*pi*fskDeviation*5)  (factories/QuadDemodFactory.h:108-110)
  [[nodiscard]] virtual Result<Filter> createQuadDemod(Modulation modulation, float rfSampleRate, float fskDeviation,
                                                       ICudaCommandQueue* commandQueue) noexcept = 0;
  ABSTRACT_IREF(IQuadDemodFactory);
};
class IRfToPcmAudioFactory : public INodeFactory {
 public:
  // In this library the returned Filter is ONE node running the fused kernels (mix -> FIR -> demod -> audio FIR),
  // not the five-node Component the reference assembles (factories/RfToPcmAudioFactory.cpp:214-304).
  [[nodiscard]] virtual Result<Filter> createRfToPcm(float rfSampleRate, Modulation modulation, size_t rfLowPassDecim,
                                                     size_t audioLowPassDecim, float centerFrequency, float channelFrequency,
                                                     float channelWidth, float fskDeviationIfFm, float rfLowPassDbAttenuation,
                                                     float audioLowPassDbAttenuation, const char* commandQueueId) noexcept = 0;
  ABSTRACT_IREF(IRfToPcmAudioFactory);
};
class IReadByteCountMonitorFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IReadByteCountMonitor> create(Filter* monitoredFilter) noexcept = 0;
  ABSTRACT_IREF(IReadByteCountMonitorFactory);
};

#endif  // GPUSDRPIPELINE_ABI_NODES_H
