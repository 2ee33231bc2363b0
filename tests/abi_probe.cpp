// Compile-time and no-GPU run-time checks of the gpusdrpipeline boundary (built and run by tests/test_host_abi.py).
// Layout facts are the ones measured on the reference's headers in SURVEY.md section 8(b).
#include <gpusdrpipeline/Factories.h>
#include <gpusdrpipeline/FusedChain.h>
#include <gpusdrpipeline/filters/BaseFilter.h>

#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

static_assert(sizeof(Status) == 4 && Status_ParseError == 9, "Status is uint32_t with ten codes in fixed order");
static_assert(SampleType_FloatComplex == 0 && SampleType_Float == 1 && SampleType_Int8Complex == 2, "SampleType values");
static_assert(Modulation_Am == 0 && Modulation_Fm == 1, "Modulation values");
static_assert(sizeof(RefResult<IBuffer>) == 16 && offsetof(RefResult<IBuffer>, value) == 8, "RefResult layout");
static_assert(sizeof(ValResult<size_t>) == 16 && offsetof(ValResult<size_t>, value) == 8, "ValResult<size_t> layout");
static_assert(sizeof(ValResult<int32_t>) == 8 && offsetof(ValResult<int32_t>, value) == 4, "ValResult<int32_t> layout");
static_assert(std::is_trivially_copyable<RefResult<IBuffer>>::value, "results are returned in registers");
static_assert(std::is_same<Result<IBuffer>, RefResult<IBuffer>>::value && std::is_same<Result<int>, ValResult<int>>::value, "Result selection");
static_assert(std::is_base_of<Sink, Filter>::value && std::is_base_of<Source, Filter>::value && std::is_base_of<Node, IDriver>::value, "hierarchy");
static_assert(std::is_base_of<IDriver, IFilterDriver>::value && std::is_base_of<Filter, IFilterDriver>::value, "IFilterDriver is a driver and a Filter");

// host end of the host-only graph below: appends whatever the driver hands it to a system-memory buffer
class CollectSink final : public Sink {
 public:
  CollectSink(IBuffer* storage, IBufferSliceFactory* slices) : mStorage(storage), mSlices(slices) {}
  Result<IBuffer> requestBuffer(size_t, size_t byteCount) noexcept final {
    if (mStorage->range()->remaining() < byteCount) return ERR_RESULT(Status_OutOfMemory);
    return mSlices->sliceRemaining(mStorage.get());
  }
  Status commitBuffer(size_t, size_t byteCount) noexcept final { return mStorage->range()->increaseEndOffset(byteCount); }
  size_t preferredInputBufferSize(size_t) noexcept final { return size_t(1) << 16; }

 private:
  ConstRef<IBuffer> mStorage;
  ConstRef<IBufferSliceFactory> mSlices;
  REF_COUNTED(CollectSink);
};

// an out-of-tree Filter on the published helper base class with SYSTEM-memory port buffers: copies its input through
class HostPassThrough final : public BaseFilter {
 public:
  static Ref<Filter> create(IFactories* f) {
    ConstRef<IRelocatableResizableBufferFactory> buffers = unwrap(f->createRelocatableSysMemBufferFactory());
    std::vector<ImmutableRef<IBufferCopier>> copiers;
    copiers.emplace_back(f->getSysMemCopier());
    return Ref<Filter>(new HostPassThrough(buffers.get(), f->getBufferSliceFactory(), std::move(copiers), f->getSysMemCopier()));
  }
  size_t getOutputDataSize(size_t port) noexcept final {
    if (port != 0) return 0;
    Result<IBuffer> in = getPortInputBuffer(0);
    if (in.status != Status_Success) return 0;
    ConstRef<IBuffer> hold(in.value);
    return hold->range()->used();
  }
  size_t getOutputSizeAlignment(size_t) noexcept final { return 1; }
  size_t preferredInputBufferSize(size_t) noexcept final { return size_t(1) << 16; }
  Status readOutput(IBuffer** bufs, size_t) noexcept final {
    Ref<IBuffer> in;
    UNWRAP_OR_FWD_STATUS(in, getPortInputBuffer(0));
    size_t n = in->range()->used();
    if (n > bufs[0]->range()->remaining()) n = bufs[0]->range()->remaining();
    FWD_IF_ERR(mCopier->copy(bufs[0]->writePtr(), in->readPtr(), n));
    FWD_IF_ERR(bufs[0]->range()->increaseEndOffset(n));
    return consumeInputBytesAndMoveUsedToStart(0, n);
  }

 private:
  HostPassThrough(IRelocatableResizableBufferFactory* buffers, IBufferSliceFactory* slices, std::vector<ImmutableRef<IBufferCopier>>&& copiers,
                  IBufferCopier* copier)
      : BaseFilter(buffers, slices, 1, std::move(copiers)), mCopier(copier) {}
  ConstRef<IBufferCopier> mCopier;
  REF_COUNTED(HostPassThrough);
};

int main(int argc, char** argv) {
  gslogSetVerbosity(GSLOG_FATAL);
  Result<IFactories> r = getFactoriesSingleton();
  if (r.status != Status_Success || r.value == nullptr) return 1;
  IFactories* f = r.value;
  const void* getters[] = {
      f->getResizableBufferFactory(), f->getCudaAllocatorFactory(), f->getBufferSliceFactory(), f->getSysMemAllocator(), f->getSysMemCopier(),
      f->getCudaBufferCopierFactory(), f->getBufferUtil(), f->getCudaMemcpyFilterFactory(), f->getAacFileWriterFactory(), f->getAddConstFactory(),
      f->getAddConstToVectorLengthFactory(), f->getCosineSourceFactory(), f->getFileReaderFactory(), f->getFirFactory(),
      f->getHackrfSourceFactory(), f->getInt8ToFloatFactory(), f->getMagnitudeFactory(), f->getMultiplyFactory(), f->getQuadDemodFactory(),
      f->getSysMemSet(), f->getCudaMemSetFactory(), f->getSteppingDriverFactory(), f->getFilterDriverFactory(),
      f->getPortRemappingSinkFactory(), f->getPortRemappingSourceFactory(), f->getRfToPcmAudioFactory(), f->getReadByteCountMonitorFactory(),
      f->getDriverToDotFactory(), f->getBufferRangeFactory(), f->getCommandQueueFactory(), f->getCudaCommandQueueFactory()};
  for (const void* g : getters)
    if (g == nullptr) return 2;
  if (getFactoriesSingleton().value != f) return 3;

  // host-only pieces work without a GPU: system-memory buffers, ranges, slices, pools
  ConstRef<IBufferFactory> buffers = unwrap(f->createSysMemBufferFactory());
  ConstRef<IBuffer> buffer = unwrap(buffers->createBuffer(100));
  if (buffer->range()->capacity() != 100 || buffer->range()->used() != 0) return 4;
  if (buffer->range()->setUsedRange(10, 60) != Status_Success || buffer->range()->setUsedRange(70, 60) == Status_Success) return 5;
  ConstRef<IBuffer> slice = unwrap(f->getBufferSliceFactory()->slice(buffer, 40, 80));
  if (slice->range()->capacity() != 40 || slice->range()->offset() != 0 || slice->range()->endOffset() != 20) return 6;
  if (slice->base() != buffer->base() + 40) return 7;
  ConstRef<IBuffer> rest = unwrap(f->getBufferSliceFactory()->sliceRemaining(buffer));
  if (rest->range()->capacity() != 40 || rest->range()->used() != 0 || rest->base() != buffer->base() + 60) return 8;
  ConstRef<IBufferPool> pool = unwrap(f->createBufferPool(2, 64, buffers));
  {
    ConstRef<IBuffer> a = unwrap(pool->getBuffer());
    ConstRef<IBuffer> b = unwrap(pool->getBuffer());
    Result<IBuffer> none = pool->tryGetBuffer();
    if (none.status != Status_Success || none.value != nullptr) return 9;
  }
  ConstRef<IBuffer> again = unwrap(pool->tryGetBuffer());
  if (again == nullptr) return 10;
  ConstRef<IRelocatableResizableBufferFactory> relocFactory = unwrap(f->createRelocatableSysMemBufferFactory());
  ConstRef<IRelocatableResizableBuffer> reloc = unwrap(relocFactory->createRelocatableBuffer(64));
  for (int i = 0; i < 64; i++) reloc->base()[i] = static_cast<uint8_t>(i);
  if (reloc->range()->setUsedRange(40, 64) != Status_Success || reloc->relocateUsedToStart() != Status_Success) return 11;
  if (reloc->range()->offset() != 0 || reloc->range()->used() != 24 || reloc->base()[0] != 40 || reloc->base()[23] != 63) return 12;
  if (reloc->range()->setUsedRange(4, 64) != Status_Success || reloc->relocateUsedToStart() != Status_Success || reloc->base()[59] != 63 + 0) return 13;
  if (reloc->resize(1000) != Status_Success || reloc->range()->capacity() < 1000 || reloc->base()[0] != 44) return 14;

  // the registry knows the reference's node names; unknown names and bad JSON are errors, not crashes
  if (!hasNodeFactory("Fir") || !hasNodeFactory("MultiplyCCC") || !hasNodeFactory("Component") || hasNodeFactory("nope")) return 15;
  if (createNode("nope", "{}").status != Status_NotFound) return 16;
  if (createNode("Fir", "{not json").status != Status_ParseError) return 17;
  if (f->getHackrfSourceFactory()->createHackrfSource(0, 0, 1.0, 3).status != Status_NotFound) return 18;
  if (f->getCommandQueueFactory()->exists("q0")) return 19;
  if (f->getCommandQueueFactory()->getCudaCommandQueue("q0").status != Status_NotFound) return 20;

  // "Component" descriptions in the reference's schema (FilterDriverFactory.cpp:27-179) parse without a GPU as long as no
  // GPU node has to be built: nodes as an OBJECT keyed by id, exposedPort/mapped{node,port} lists, the documented statuses
  {
    Result<Node> empty = createNode("Component", "{\"nodes\": {}, \"connections\": [], \"inputPorts\": [], \"outputPorts\": []}");
    if (empty.status != Status_Success || empty.value == nullptr) return 30;
    ConstRef<Node> hold(empty.value);
    if (hold->asFilter() == nullptr || hold->asDriver() == nullptr) return 31;
  }
  if (createNode("Component", "{\"connections\": []}").status != Status_ParseError) return 32;  // "nodes" is required
  if (createNode("Component", "{\"nodes\": {\"a\": {\"description\": \"no type\"}}}").status != Status_InvalidArgument) return 33;
  if (createNode("Component", "{\"nodes\": {\"a\": {\"type\": \"nope\"}}}").status != Status_NotFound) return 34;
  if (createNode("Component", "{\"nodes\": {}, \"inputPorts\": [{\"exposedPort\": 0, \"mapped\": {\"node\": \"ghost\", \"port\": 1}}]}").status !=
      Status_NotFound)
    return 35;
  if (createNode("Component", "{\"nodes\": {}, \"connections\": [{\"source\": \"a\", \"sink\": \"b\"}]}").status != Status_InvalidArgument) return 36;
  // a nested Component is a node like any other (ids are per Component)
  if (createNode("Component", "{\"nodes\": {\"inner\": {\"type\": \"Component\", \"nodes\": {}}}, \"outputPort\": \"inner\"}").status != Status_Success)
    return 37;
  // the reference's registry names (FilterFactories.cpp:132-150) all resolve
  for (const char* name : {"AacWriter", "AddConst", "AddConstToVectorLength", "Component", "Cosine", "File", "Fir", "HackRfSource", "Int8ToFloat",
                           "Magnitude", "MultiplyCCC", "QuadDemod"})
    if (!hasNodeFactory(name)) return 38;

  // A graph of host-only nodes runs without a GPU: FileReader -> ReadByteCountMonitor(BaseFilter pass-through) -> host sink under ISteppingDriver::connect
  // + doFilter() (SteppingDriver.cpp:193-366); every byte arrives once and in order, the monitor counts them, DriverToDot names
  // the nodes.  argv[1] = a scratch file this block writes.
  if (argc > 1) {
    std::vector<uint8_t> bytes(200000 + 37);
    for (size_t i = 0; i < bytes.size(); i++) bytes[i] = static_cast<uint8_t>((i * 2654435761u) >> 13);
    FILE* out = fopen(argv[1], "wb");
    if (out == nullptr || fwrite(bytes.data(), 1, bytes.size(), out) != bytes.size()) return 40;
    fclose(out);
    ConstRef<Source> reader = unwrap(f->getFileReaderFactory()->createFileReader(argv[1]));
    Ref<Filter> through = HostPassThrough::create(f);
    ConstRef<IReadByteCountMonitor> monitor = unwrap(f->getReadByteCountMonitorFactory()->create(through.get()));
    ConstRef<IBuffer> storage = unwrap(buffers->createBuffer(bytes.size() + (size_t(1) << 17)));
    ConstRef<CollectSink> sink(new CollectSink(storage, f->getBufferSliceFactory()));
    Ref<ISteppingDriver> driver = unwrap(f->getSteppingDriverFactory()->createSteppingDriver());
    if (driver->connect(reader.get(), 0, monitor.get(), 0) != Status_Success || driver->connect(monitor.get(), 0, sink.get(), 0) != Status_Success)
      return 41;
    if (driver->setupNode(reader.get(), "Read the file") != Status_Success || driver->setupNode(monitor.get(), "Pass through, counted") != Status_Success ||
        driver->setupNode(sink.get(), "Collect") != Status_Success)
      return 42;
    size_t steps = 0;
    while (monitor->getByteCountRead(0) < bytes.size() && steps < 1000) {
      if (driver->doFilter() != Status_Success) return 43;
      steps++;
    }
    if (steps < 3 || monitor->getByteCountRead(0) != bytes.size() || storage->range()->used() != bytes.size()) return 44;  // 64 KiB per pass
    if (std::memcmp(storage->readPtr(), bytes.data(), bytes.size()) != 0) return 45;
    if (driver->doFilter() != Status_Success || storage->range()->used() != bytes.size()) return 46;  // end of file: nothing more, no error
    ConstRef<IDriverToDiagram> toDot = unwrap(f->getDriverToDotFactory()->create());
    const size_t need = unwrap(toDot->convertToDot(driver.get(), "host_graph", nullptr, 0));
    std::string text(need, '\0');
    (void)unwrap(toDot->convertToDot(driver.get(), "host_graph", text.data(), text.size()));
    if (text.find("digraph") != 0 || text.find("Read the file") == std::string::npos || text.find("->") == std::string::npos) return 47;
    if (createNode("File", "{\"fileName\": \"/nonexistent/b200sdr\"}").status == Status_Success) return 48;

    // the same source inside a "Component" (reference schema, FilterDriverFactory.cpp:27-179): a "File" node exposed as output
    // port 0 through PortRemappingSource; the Component is the Source of the outer stepping driver
    const std::string json = std::string("{\"nodes\": {\"file\": {\"type\": \"File\", \"fileName\": \"") + argv[1] +
                             "\"}}, \"connections\": [], \"outputPorts\": [{\"exposedPort\": 0, \"mapped\": {\"node\": \"file\", \"port\": 0}}]}";
    Result<Node> made = createNode("Component", json.c_str());
    if (made.status != Status_Success || made.value == nullptr) return 49;
    ConstRef<Node> component(made.value);
    if (component->asSource() == nullptr) return 50;
    ConstRef<IBuffer> storage2 = unwrap(buffers->createBuffer(bytes.size() + (size_t(1) << 17)));
    ConstRef<CollectSink> sink2(new CollectSink(storage2, f->getBufferSliceFactory()));
    Ref<ISteppingDriver> outer = unwrap(f->getSteppingDriverFactory()->createSteppingDriver());
    if (outer->connect(component->asSource(), 0, sink2.get(), 0) != Status_Success) return 51;
    for (size_t pass = 0; pass < 1000 && storage2->range()->used() < bytes.size(); pass++)
      if (outer->doFilter() != Status_Success) return 52;
    if (storage2->range()->used() != bytes.size() || std::memcmp(storage2->readPtr(), bytes.data(), bytes.size()) != 0) return 53;
  }

  // no CPU fallback: with no CUDA device every GPU-facing creation fails with a Status, never with a crash
  int deviceCount = 0;
  if (cudaGetDeviceCount(&deviceCount) != cudaSuccess || deviceCount == 0) {
    if (f->getCudaCommandQueueFactory()->create(0).status == Status_Success) return 21;
  }
  std::printf("abi_probe ok (%d CUDA devices)\n", deviceCount);
  return 0;
}
