// graph_probe.cpp -- drives GRAPHS through the reference-facing host API (getFactoriesSingleton(), ISteppingDriver,
// IFilterDriver, createFilter("Component", json), PortRemappingSink/Source, ReadByteCountMonitor, DriverToDot), the way
// the reference's applications do (src/applications/am_test.cpp:295-494): nested FilterDrivers joined by one
// SteppingDriver whose doFilter() is called until a ReadByteCountMonitor has seen the expected number of bytes.
// TEST INFRASTRUCTURE: built and run by tests/test_gpu_graph.py, which compares the audio with the CPU oracle.
//
//   --mode stepping    host int8 -> [H2D -> Int8ToFloat] -> [[cosine -> Multiply <- port 0] -> Fir] -> QuadDemod -> Fir]
//                      -> [monitor(D2H) -> host]; every [..] is a FilterDriver (am_test.cpp:325-433)
//   --mode component   the same, with the RF-to-audio part built by createFilter("Component", json) in the reference's
//                      schema (FilterDriverFactory.cpp:27-179) -- the exposed input port is mapped onto Multiply port 1 and
//                      the cosine onto port 0 (a NON-identity mapping), the output through a PortRemappingSource
//                      (--input cf32: complex-float samples from the host, no Int8ToFloat node;
//                       --source file: the FileReader node of the library reads --in itself)
//   --mode elementwise host cf32 -> H2D -> PassThrough (an out-of-tree filter derived from BaseFilter) ->
//                      AddConstToVectorLength -> Magnitude -> AddConst -> monitor(D2H) -> host
// Prints one JSON line; --chunks FILE receives the element count of every cosine readOutput (the reference's float32
// phase bookkeeping depends on it, CosineSource.cpp:51,72,82); --dot FILE the Graphviz text of the outer driver.
#include <cuComplex.h>
#include <cuda_runtime.h>
#include <gpusdrpipeline/Factories.h>
#include <gpusdrpipeline/filters/BaseFilter.h>

#include <time.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

using namespace std;

static vector<char> readFile(const string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    fprintf(stderr, "cannot open %s\n", path.c_str());
    exit(2);
  }
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  vector<char> data(static_cast<size_t>(n));
  if (n > 0 && fread(data.data(), 1, data.size(), f) != data.size()) exit(2);
  fclose(f);
  return data;
}
static vector<float> readFloats(const string& path) {
  const vector<char> raw = readFile(path);
  vector<float> v(raw.size() / sizeof(float));
  memcpy(v.data(), raw.data(), v.size() * sizeof(float));
  return v;
}

// ---- host-side ends of the graph (stand-ins for HackrfSource / AacFileWriter) ----------------------------------------
class HostSource final : public Source {
 public:
  HostSource(const vector<char>& data, size_t chunk, IBufferCopier* copier) : mData(data), mChunk(chunk), mCopier(copier) {}
  size_t getOutputDataSize(size_t port) noexcept final {
    if (port != 0) return 0;
    const size_t left = mData.size() - mPos;
    return left < mChunk ? left : mChunk;  // like one HackRF transfer (HackrfSource.cpp:175-261)
  }
  size_t getOutputSizeAlignment(size_t) noexcept final { return 1; }
  IBufferCopier* getOutputCopier(size_t) noexcept final { return mCopier.get(); }
  Status readOutput(IBuffer** bufs, size_t) noexcept final {
    size_t n = getOutputDataSize(0);
    if (n > bufs[0]->range()->remaining()) n = bufs[0]->range()->remaining();
    memcpy(bufs[0]->writePtr(), mData.data() + mPos, n);
    mPos += n;
    return bufs[0]->range()->increaseEndOffset(n);
  }
  bool exhausted() const { return mPos == mData.size(); }

 private:
  const vector<char>& mData;
  const size_t mChunk;
  size_t mPos = 0;
  ConstRef<IBufferCopier> mCopier;
  REF_COUNTED(HostSource);
};

class HostSink final : public Sink {
 public:
  HostSink(IBuffer* storage, IBufferSliceFactory* slices) : mStorage(storage), mSlices(slices) {}
  Result<IBuffer> requestBuffer(size_t, size_t byteCount) noexcept final {
    if (mStorage->range()->remaining() < byteCount) return ERR_RESULT(Status_OutOfMemory);
    return mSlices->sliceRemaining(mStorage.get());
  }
  Status commitBuffer(size_t, size_t byteCount) noexcept final { return mStorage->range()->increaseEndOffset(byteCount); }
  size_t preferredInputBufferSize(size_t) noexcept final { return size_t(1) << 20; }
  size_t bytes() const { return mStorage->range()->used(); }
  const uint8_t* data() const { return mStorage->readPtr(); }

 private:
  ConstRef<IBuffer> mStorage;
  ConstRef<IBufferSliceFactory> mSlices;
  REF_COUNTED(HostSink);
};

// logs how many elements each readOutput produced (the cosine's float32 phase depends on the chunking)
class LoggingSource final : public Source {
 public:
  LoggingSource(Source* inner, size_t elemBytes, vector<size_t>* log) : mInner(inner), mElemBytes(elemBytes), mLog(log) {}
  size_t getOutputDataSize(size_t port) noexcept final { return mInner->getOutputDataSize(port); }
  size_t getOutputSizeAlignment(size_t port) noexcept final { return mInner->getOutputSizeAlignment(port); }
  IBufferCopier* getOutputCopier(size_t port) noexcept final { return mInner->getOutputCopier(port); }
  Status readOutput(IBuffer** bufs, size_t n) noexcept final {
    const size_t before = bufs[0]->range()->endOffset();
    const Status st = mInner->readOutput(bufs, n);
    mLog->push_back((bufs[0]->range()->endOffset() - before) / mElemBytes);
    return st;
  }

 private:
  ConstRef<Source> mInner;
  const size_t mElemBytes;
  vector<size_t>* const mLog;
  REF_COUNTED(LoggingSource);
};

// An out-of-tree filter written against the published helper base class (filters/BaseFilter.h): copies its input through.
class PassThrough final : public BaseFilter {
 public:
  static Ref<Filter> create(IFactories* f, ICudaCommandQueue* queue) {
    ConstRef<IRelocatableResizableBufferFactory> buffers = unwrap(f->createRelocatableCudaBufferFactory(queue, 32, false));
    ConstRef<IBufferCopier> d2d = unwrap(f->getCudaBufferCopierFactory()->createBufferCopier(queue, cudaMemcpyDeviceToDevice));
    vector<ImmutableRef<IBufferCopier>> copiers;
    copiers.emplace_back(d2d.get());
    return Ref<Filter>(new PassThrough(buffers.get(), f->getBufferSliceFactory(), std::move(copiers), d2d.get()));
  }
  size_t getOutputDataSize(size_t port) noexcept final {
    if (port != 0) return 0;
    Result<IBuffer> in = getPortInputBuffer(0);
    if (in.status != Status_Success) return 0;
    ConstRef<IBuffer> hold(in.value);
    return hold->range()->used();
  }
  size_t getOutputSizeAlignment(size_t) noexcept final { return 8; }
  size_t preferredInputBufferSize(size_t) noexcept final { return size_t(1) << 20; }
  Status readOutput(IBuffer** bufs, size_t) noexcept final {
    Ref<IBuffer> in;
    UNWRAP_OR_FWD_STATUS(in, getPortInputBuffer(0));
    size_t n = in->range()->used();
    if (n > bufs[0]->range()->remaining()) n = bufs[0]->range()->remaining();
    n -= n % 8;
    FWD_IF_ERR(mCopier->copy(bufs[0]->writePtr(), in->readPtr(), n));
    FWD_IF_ERR(bufs[0]->range()->increaseEndOffset(n));
    return consumeInputBytesAndMoveUsedToStart(0, n);
  }

 private:
  PassThrough(IRelocatableResizableBufferFactory* buffers, IBufferSliceFactory* slices, vector<ImmutableRef<IBufferCopier>>&& copiers,
              IBufferCopier* copier)
      : BaseFilter(buffers, slices, 1, std::move(copiers)), mCopier(copier) {}
  ConstRef<IBufferCopier> mCopier;
  REF_COUNTED(PassThrough);
};

// registerNodeFactory(): a user-registered node type for Component descriptions -- the library's "Cosine" behind the logger
class LoggedCosineFactory final : public INodeFactory {
 public:
  explicit LoggedCosineFactory(vector<size_t>* log) : mLog(log) {}
  Result<Node> create(const char* json) noexcept final {
    Result<Node> inner = createNode("Cosine", json);
    if (inner.status != Status_Success) return inner;
    ConstRef<Node> hold(inner.value);
    return makeRefResultNonNull<Node>(new (std::nothrow) LoggingSource(hold->asSource(), sizeof(cuComplex), mLog));
  }

 private:
  vector<size_t>* const mLog;
  REF_COUNTED(LoggedCosineFactory);
};

// --mode memcpy: the pinned input port of a host->device CudaMemcpyFilter under partial drains.  The stream is stalled
// by a host callback, so every copy is still queued when the host asks for the next buffer: the node must not hand out
// (or compact over) bytes a queued copy has yet to read.
static int memcpyMode(IFactories* f, ICudaCommandQueue* queue) {
  const size_t MiB = size_t(1) << 20, piece = 256 << 10;
  auto pattern = [](size_t i, unsigned salt) { return static_cast<uint8_t>((i * 131u + 7u + salt * 29u) ^ (i >> 11)); };
  ConstRef<Filter> h2d = unwrap(f->getCudaMemcpyFilterFactory()->createCudaMemcpy(cudaMemcpyHostToDevice, queue));
  ConstRef<IAllocator> deviceAlloc = unwrap(f->getCudaAllocatorFactory()->createCudaAllocator(queue, 256, false));
  ConstRef<IBufferFactory> deviceBuffers = unwrap(f->createBufferFactory(deviceAlloc));
  ConstRef<IBuffer> device = unwrap(deviceBuffers->createBuffer(6 * MiB));
  size_t devicePos = 0;
  auto drain = [&](size_t bytes) {  // readOutput into a view of `piece` bytes at a time
    for (size_t done = 0; done < bytes; done += piece) {
      ConstRef<IBuffer> view = unwrap(f->getBufferSliceFactory()->slice(device, devicePos, devicePos + piece));
      view->range()->clearRange();
      IBuffer* out[1] = {view.get()};
      THROW_IF_ERR(h2d->readOutput(out, 1));
      if (view->range()->used() != piece) throw runtime_error("short read from the copy node");
      devicePos += piece;
    }
  };
  vector<uint8_t> expect;
  auto feed = [&](size_t bytes, unsigned salt) {
    Ref<IBuffer> staged = unwrap(h2d->requestBuffer(0, bytes));
    if (staged->range()->remaining() < bytes) throw runtime_error("requestBuffer returned too little room");
    for (size_t i = 0; i < bytes; i++) {
      staged->writePtr()[i] = pattern(i, salt);
      expect.push_back(pattern(i, salt));
    }
    THROW_IF_ERR(h2d->commitBuffer(0, bytes));
  };
  cudaSetDevice(queue->cudaDevice());
  feed(3 * MiB, 1);
  cudaLaunchHostFunc(queue->cudaStream(), [](void*) { struct timespec ts = {0, 150000000}; nanosleep(&ts, nullptr); }, nullptr);
  drain(2 * MiB);       // 8 partial drains, all queued behind the stall: 1 MiB is left at offset 2 MiB of the pinned block
  feed(1 * MiB, 2);     // fits only after compaction to the front -- over bytes the queued copies have not read yet
  drain(MiB + MiB / 2);
  feed(2 * MiB, 3);     // behind the tail / growth while copies are pending
  drain(2 * MiB + MiB / 2);
  if (h2d->getOutputDataSize(0) != 0) return 3;
  vector<uint8_t> got(devicePos);
  cudaMemcpyAsync(got.data(), device->base(), devicePos, cudaMemcpyDeviceToHost, queue->cudaStream());
  cudaStreamSynchronize(queue->cudaStream());
  size_t bad = 0;
  for (size_t i = 0; i < got.size(); i++) bad += got[i] != expect[i];
  printf("{\"mode\": \"memcpy\", \"bytes\": %zu, \"mismatches\": %zu}\n", got.size(), bad);
  return bad == 0 && got.size() == expect.size() ? 0 : 4;
}

struct Args {
  string mode = "stepping", in, out, taps1, taps2, mod = "am", dot, chunks, input = "int8", source = "host";
  double fs = 19.2e6, freq = 0, dev = 75e3, addMag = 0.25, addConst = -0.125;
  size_t d1 = 1, d2 = 1, chunk = 262144;
};

static string num(double v) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.9g", v);
  return buf;
}
static string tapList(const vector<float>& taps) {
  string s = "[";
  for (size_t i = 0; i < taps.size(); i++) s += (i ? "," : "") + num(taps[i]);
  return s + "]";
}

int main(int argc, char** argv) {
  Args a;
  for (int i = 1; i + 1 < argc; i += 2) {
    const string k = argv[i], v = argv[i + 1];
    if (k == "--mode") a.mode = v;
    else if (k == "--in") a.in = v;
    else if (k == "--out") a.out = v;
    else if (k == "--taps1") a.taps1 = v;
    else if (k == "--taps2") a.taps2 = v;
    else if (k == "--mod") a.mod = v;
    else if (k == "--dot") a.dot = v;
    else if (k == "--chunks") a.chunks = v;
    else if (k == "--input") a.input = v;
    else if (k == "--source") a.source = v;
    else if (k == "--fs") a.fs = atof(v.c_str());
    else if (k == "--freq") a.freq = atof(v.c_str());
    else if (k == "--dev") a.dev = atof(v.c_str());
    else if (k == "--d1") a.d1 = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--d2") a.d2 = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--chunk") a.chunk = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--add-mag") a.addMag = atof(v.c_str());
    else if (k == "--add-const") a.addConst = atof(v.c_str());
    else {
      fprintf(stderr, "unknown argument %s\n", k.c_str());
      return 2;
    }
  }
  gslogSetVerbosity(GSLOG_WARN);
  if (a.mode == "memcpy") {
    ConstRef<IFactories> factories = unwrap(getFactoriesSingleton());
    ConstRef<ICudaCommandQueue> q = unwrap(factories->getCudaCommandQueueFactory()->create(0));
    return memcpyMode(factories, q);
  }
  const vector<char> input = readFile(a.in);
  const bool fm = a.mod == "fm";
  const bool chainMode = a.mode == "stepping" || a.mode == "component";
  const vector<float> taps1 = chainMode ? readFloats(a.taps1) : vector<float>(), taps2 = chainMode ? readFloats(a.taps2) : vector<float>();

  ConstRef<IFactories> f = unwrap(getFactoriesSingleton());
  THROW_IF_ERR(f->getCommandQueueFactory()->create("q0", "{\"queueType\": \"cuda\", \"cudaDevice\": 0}"));
  ConstRef<ICudaCommandQueue> queue = unwrap(f->getCommandQueueFactory()->getCudaCommandQueue("q0"));
  const float rfRate = static_cast<float>(a.fs), demodRate = static_cast<float>(a.fs / static_cast<double>(a.d1));
  vector<size_t> cosineChunks;

  // ---- input pipeline: FilterDriver used as a Source (am_test.cpp:352-377) ------------------------------------------
  ConstRef<HostSource> hostSource(new HostSource(input, a.chunk, f->getSysMemCopier()));
  ConstRef<Filter> h2d = unwrap(f->getCudaMemcpyFilterFactory()->createCudaMemcpy(cudaMemcpyHostToDevice, queue));
  Ref<IFilterDriver> inputPipeline = unwrap(f->getFilterDriverFactory()->createFilterDriver());
  // --source file: the samples come from the library's own FileReader node (FileReader.cpp:48-67) instead of the stand-in above
  Ref<Source> fileSource;
  Source* sourceNode = hostSource.get();
  if (a.source == "file") {
    fileSource = unwrap(f->getFileReaderFactory()->createFileReader(a.in.c_str()));
    sourceNode = fileSource.get();
  }
  THROW_IF_ERR(inputPipeline->connect(sourceNode, 0, h2d, 0));
  THROW_IF_ERR(inputPipeline->setupNode(sourceNode, a.source == "file" ? "Read samples from a file" : "Host samples"));
  THROW_IF_ERR(inputPipeline->setupNode(h2d, "Copy samples to GPU memory"));
  Ref<Filter> middle;
  size_t outElemBytes = 4;
  if (chainMode) {
    if (a.input == "int8") {
      ConstRef<Filter> int8ToFloat = unwrap(f->getInt8ToFloatFactory()->createFilter(queue));
      THROW_IF_ERR(inputPipeline->connect(h2d, 0, int8ToFloat, 0));
      THROW_IF_ERR(inputPipeline->setupNode(int8ToFloat, "Convert complex int8 to complex float"));
      inputPipeline->setDriverOutput(int8ToFloat);
    } else {
      // complex-float samples straight from the host.  (The reference's own Int8ToFloat rejects the one-port readOutput its
      // SteppingDriver issues -- Int8ToFloat.cpp:81 tests `0 == portCount` -- so the reference framework is driven this way.)
      inputPipeline->setDriverOutput(h2d);
    }
    if (a.mode == "stepping") {
      // frequency shifter: a FilterDriver whose input is a PortRemappingSink exposing Multiply port 0 (am_test.cpp:295-350)
      ConstRef<Source> cosineRaw = unwrap(f->getCosineSourceFactory()->createCosineSource(SampleType_FloatComplex, rfRate, static_cast<float>(a.freq), queue));
      ConstRef<LoggingSource> cosine(new LoggingSource(cosineRaw, sizeof(cuComplex), &cosineChunks));
      ConstRef<Filter> multiply = unwrap(f->getMultiplyFactory()->createFilter(queue));
      ConstRef<Filter> rfFir = unwrap(f->getFirFactory()->createFir(SampleType_Float, SampleType_FloatComplex, a.d1, taps1.data(), taps1.size(), queue));
      ConstRef<IPortRemappingSink> port0 = unwrap(f->getPortRemappingSinkFactory()->create());
      port0->addPortMapping(0, multiply, 0);
      Ref<IFilterDriver> shifter = unwrap(f->getFilterDriverFactory()->createFilterDriver());
      THROW_IF_ERR(shifter->setupNode(cosine.get(), "Produce a cosine signal"));
      THROW_IF_ERR(shifter->setupNode(multiply, "Multiply signals"));
      THROW_IF_ERR(shifter->setupNode(rfFir, "Low-pass filter"));
      shifter->setDriverInput(port0);
      shifter->setDriverOutput(rfFir);
      THROW_IF_ERR(shifter->connect(cosine.get(), 0, multiply, 1));
      THROW_IF_ERR(shifter->connect(multiply, 0, rfFir, 0));
      // RF -> audio: shifter -> demodulator -> audio filter (am_test.cpp:379-433)
      ConstRef<Filter> demod = unwrap(f->getQuadDemodFactory()->createQuadDemod(fm ? Modulation_Fm : Modulation_Am, demodRate, static_cast<float>(a.dev), queue));
      ConstRef<Filter> audioFir = unwrap(f->getFirFactory()->createFir(SampleType_Float, SampleType_Float, a.d2, taps2.data(), taps2.size(), queue));
      Ref<IFilterDriver> rfToAudio = unwrap(f->getFilterDriverFactory()->createFilterDriver());
      rfToAudio->setDriverInput(shifter.get());
      rfToAudio->setDriverOutput(audioFir);
      THROW_IF_ERR(rfToAudio->connect(shifter.get(), 0, demod, 0));
      THROW_IF_ERR(rfToAudio->connect(demod, 0, audioFir, 0));
      THROW_IF_ERR(rfToAudio->setupNode(shifter.get(), "Shift RF frequency of channel"));
      THROW_IF_ERR(rfToAudio->setupNode(demod, "Demodulate"));
      THROW_IF_ERR(rfToAudio->setupNode(audioFir, "Process audio"));
      middle = rfToAudio.get();
    } else {
      ConstRef<LoggedCosineFactory> loggedCosine(new LoggedCosineFactory(&cosineChunks));
      THROW_IF_ERR(registerNodeFactory("LoggedCosine", loggedCosine.get()));
      const string q = "\"commandQueue\": \"q0\"";
      const string json =
          "{\"nodes\": {"
          "\"cosineSource\": {\"type\": \"LoggedCosine\", \"description\": \"Produce a cosine signal\", \"sampleType\": \"FloatComplex\", \"sampleRate\": " +
          num(rfRate) + ", \"frequency\": " + num(static_cast<float>(a.freq)) + ", " + q +
          "}, \"multiplyForFrequencyShift\": {\"type\": \"MultiplyCCC\", " + q +
          "}, \"rfLowPassFilter\": {\"type\": \"Fir\", \"taps\": " + tapList(taps1) +
          ", \"tapType\": \"Float\", \"elementType\": \"FloatComplex\", \"decimation\": " + to_string(a.d1) + ", " + q +
          "}, \"quadDemod\": {\"type\": \"QuadDemod\", \"modulation\": \"" + string(fm ? "FM" : "AM") + "\", \"sampleRate\": " + num(demodRate) +
          ", \"fskDeviation\": " + num(static_cast<float>(a.dev)) + ", " + q +
          "}, \"audioLowPassFilter\": {\"type\": \"Fir\", \"taps\": " + tapList(taps2) +
          ", \"tapType\": \"Float\", \"elementType\": \"Float\", \"decimation\": " + to_string(a.d2) + ", " + q +
          "}}, \"connections\": ["
          "{\"source\": \"cosineSource\", \"sourcePort\": 0, \"sink\": \"multiplyForFrequencyShift\", \"sinkPort\": 0},"
          "{\"source\": \"multiplyForFrequencyShift\", \"sourcePort\": 0, \"sink\": \"rfLowPassFilter\", \"sinkPort\": 0},"
          "{\"source\": \"rfLowPassFilter\", \"sourcePort\": 0, \"sink\": \"quadDemod\", \"sinkPort\": 0},"
          "{\"source\": \"quadDemod\", \"sourcePort\": 0, \"sink\": \"audioLowPassFilter\", \"sinkPort\": 0}],"
          "\"inputPorts\": [{\"exposedPort\": 0, \"mapped\": {\"node\": \"multiplyForFrequencyShift\", \"port\": 1}}],"
          "\"outputPorts\": [{\"exposedPort\": 0, \"mapped\": {\"node\": \"audioLowPassFilter\", \"port\": 0}}]}";
      middle = unwrap(createFilter("Component", json.c_str()));
    }
  } else {
    // cf32 input: PassThrough (BaseFilter) -> AddConstToVectorLength -> Magnitude -> AddConst, joined by a FilterDriver
    inputPipeline->setDriverOutput(h2d);
    Ref<Filter> pass = PassThrough::create(f, queue);
    ConstRef<Filter> addMag = unwrap(f->getAddConstToVectorLengthFactory()->createAddConstToVectorLength(static_cast<float>(a.addMag), queue));
    ConstRef<Filter> magnitude = unwrap(f->getMagnitudeFactory()->createFilter(queue));
    ConstRef<Filter> addConst = unwrap(f->getAddConstFactory()->createAddConst(static_cast<float>(a.addConst), queue));
    Ref<IFilterDriver> ops = unwrap(f->getFilterDriverFactory()->createFilterDriver());
    ops->setDriverInput(pass.get());
    ops->setDriverOutput(addConst);
    THROW_IF_ERR(ops->connect(pass.get(), 0, addMag, 0));
    THROW_IF_ERR(ops->connect(addMag, 0, magnitude, 0));
    THROW_IF_ERR(ops->connect(magnitude, 0, addConst, 0));
    THROW_IF_ERR(ops->setupNode(pass.get(), "Out-of-tree BaseFilter"));
    THROW_IF_ERR(ops->setupNode(addMag, "Add to vector length"));
    THROW_IF_ERR(ops->setupNode(magnitude, "Magnitude"));
    THROW_IF_ERR(ops->setupNode(addConst, "Add constant"));
    middle = ops.get();
  }

  // ---- output pipeline: FilterDriver used as a Sink; the byte-count monitor is the stop condition (am_test.cpp:481-494) --
  ConstRef<Filter> d2h = unwrap(f->getCudaMemcpyFilterFactory()->createCudaMemcpy(cudaMemcpyDeviceToHost, queue));
  ConstRef<IReadByteCountMonitor> monitor = unwrap(f->getReadByteCountMonitorFactory()->create(d2h));
  ConstRef<IAllocator> pinned = unwrap(f->getCudaAllocatorFactory()->createCudaAllocator(queue, 32, true));
  ConstRef<IBufferFactory> pinnedBuffers = unwrap(f->createBufferFactory(pinned));
  const size_t inElems = chainMode && a.input == "int8" ? input.size() / 2 : input.size() / 8;
  ConstRef<IBuffer> storage = unwrap(pinnedBuffers->createBuffer((chainMode ? inElems / (a.d1 * a.d2) : inElems) * outElemBytes + (4 << 20)));
  ConstRef<HostSink> hostSink(new HostSink(storage, f->getBufferSliceFactory()));
  Ref<IFilterDriver> outputPipeline = unwrap(f->getFilterDriverFactory()->createFilterDriver());
  outputPipeline->setDriverInput(monitor.get());
  THROW_IF_ERR(outputPipeline->connect(monitor.get(), 0, hostSink.get(), 0));
  THROW_IF_ERR(outputPipeline->setupNode(monitor.get(), "Copy audio to host memory"));
  THROW_IF_ERR(outputPipeline->setupNode(hostSink.get(), "Host audio"));

  Ref<ISteppingDriver> driver = unwrap(f->getSteppingDriverFactory()->createSteppingDriver());
  THROW_IF_ERR(driver->connect(inputPipeline.get(), 0, middle.get(), 0));
  THROW_IF_ERR(driver->connect(middle.get(), 0, outputPipeline.get(), 0));
  THROW_IF_ERR(driver->setupNode(inputPipeline.get(), "Input Pipeline"));
  THROW_IF_ERR(driver->setupNode(middle.get(), chainMode ? "Convert RF signal to audio" : "Element-wise operations"));
  THROW_IF_ERR(driver->setupNode(outputPipeline.get(), "Output Pipeline"));

  // expected output: the count rules over the whole stream (Fir.cpp:181-186; QuadFmDemod.cpp:76-84)
  size_t expected = inElems;
  if (chainMode) {
    auto firCount = [](size_t n, size_t T, size_t D) { return n + 1 >= T ? (n + 1 - T) / D : 0; };
    const size_t rf = firCount(inElems, taps1.size(), a.d1);
    expected = firCount(fm ? (rf ? rf - 1 : 0) : rf, taps2.size(), a.d2);
  }
  size_t steps = 0, idle = 0;
  while (monitor->getByteCountRead(0) / outElemBytes < expected && idle < 64) {
    const size_t before = monitor->getByteCountRead(0);
    THROW_IF_ERR(driver->doFilter());
    steps++;
    idle = (monitor->getByteCountRead(0) == before && (a.source == "file" || hostSource->exhausted())) ? idle + 1 : 0;
  }
  cudaSetDevice(queue->cudaDevice());
  cudaStreamSynchronize(queue->cudaStream());

  if (!a.out.empty()) {
    FILE* o = fopen(a.out.c_str(), "wb");
    if (!o || fwrite(hostSink->data(), 1, hostSink->bytes(), o) != hostSink->bytes()) return 2;
    fclose(o);
  }
  if (!a.chunks.empty()) {
    FILE* o = fopen(a.chunks.c_str(), "w");
    for (size_t n : cosineChunks) fprintf(o, "%zu\n", n);
    fclose(o);
  }
  if (!a.dot.empty()) {
    ConstRef<IDriverToDiagram> toDot = unwrap(f->getDriverToDotFactory()->create());
    const size_t need = unwrap(toDot->convertToDot(driver.get(), "graph_probe", nullptr, 0));
    vector<char> text(need + 1);
    (void)unwrap(toDot->convertToDot(driver.get(), "graph_probe", text.data(), text.size()));
    FILE* o = fopen(a.dot.c_str(), "w");
    fputs(text.data(), o);
    fclose(o);
  }
  printf("{\"mode\": \"%s\", \"elements_in\": %zu, \"expected\": %zu, \"outputs\": %zu, \"monitor_bytes\": %zu, \"steps\": %zu}\n", a.mode.c_str(),
         inElems, expected, hostSink->bytes() / outElemBytes, monitor->getByteCountRead(0), steps);
  return 0;
}
