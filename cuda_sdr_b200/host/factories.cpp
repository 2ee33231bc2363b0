// getFactoriesSingleton() and the name -> node-factory registry: the root of the drop-in boundary
// (reference include/gpusdrpipeline/Factories.h:43-119, src/Factories.cpp:63-204, src/filters/FilterFactories.cpp:23-150).
#include <map>
#include <mutex>

#include "internal.h"

namespace {

using namespace gs;

class Factories final : public IFactories {
 public:
  Factories() noexcept
      : mSysMemAllocator(newSysMemAllocator()),
        mSysMemCopier(newSysMemCopier()),
        mSysMemSet(newSysMemSet()),
        mCudaAllocatorFactory(newCudaAllocatorFactory()),
        mCudaCopierFactory(newCudaBufferCopierFactory()),
        mCudaMemSetFactory(newCudaMemSetFactory()),
        mRangeFactory(newBufferRangeFactory()),
        mSliceFactory(newBufferSliceFactory()),
        mBufferUtil(newBufferUtil()),
        mResizableFactory(newResizableBufferFactory(mSysMemAllocator, mSysMemCopier)),
        mCudaQueueFactory(newCudaCommandQueueFactory()),
        mQueueFactory(newCommandQueueFactory(this)),
        mMemcpyFactory(newCudaMemcpyFilterFactory(this)),
        mAacFactory(newAacFileWriterFactory()),
        mAddConstFactory(newAddConstFactory(this)),
        mAddToMagnitudeFactory(newAddConstToVectorLengthFactory(this)),
        mCosineFactory(newCosineSourceFactory(this)),
        mFileReaderFactory(newFileReaderFactory()),
        mFirFactory(newFirFactory(this)),
        mHackrfFactory(newHackrfSourceFactory()),
        mInt8Factory(newInt8ToFloatFactory(this)),
        mMagnitudeFactory(newMagnitudeFactory(this)),
        mMultiplyFactory(newMultiplyFactory(this)),
        mQuadDemodFactory(newQuadDemodFactory(this)),
        mSteppingFactory(newSteppingDriverFactory()),
        mFilterDriverFactory(newFilterDriverFactory(this)),
        mRemapSinkFactory(newPortRemappingSinkFactory()),
        mRemapSourceFactory(newPortRemappingSourceFactory()),
        mRfToPcmFactory(newRfToPcmAudioFactory(this)),
        mMonitorFactory(newReadByteCountMonitorFactory()),
        mDotFactory(newDriverToDotFactory()) {}

  IResizableBufferFactory* getResizableBufferFactory() noexcept final { return mResizableFactory; }
  ICudaAllocatorFactory* getCudaAllocatorFactory() noexcept final { return mCudaAllocatorFactory; }
  IBufferSliceFactory* getBufferSliceFactory() final { return mSliceFactory; }
  IAllocator* getSysMemAllocator() noexcept final { return mSysMemAllocator; }
  IBufferCopier* getSysMemCopier() noexcept final { return mSysMemCopier; }
  ICudaBufferCopierFactory* getCudaBufferCopierFactory() noexcept final { return mCudaCopierFactory; }
  IBufferUtil* getBufferUtil() noexcept final { return mBufferUtil; }
  ICudaMemcpyFilterFactory* getCudaMemcpyFilterFactory() noexcept final { return mMemcpyFactory; }
  IAacFileWriterFactory* getAacFileWriterFactory() noexcept final { return mAacFactory; }
  IAddConstFactory* getAddConstFactory() noexcept final { return mAddConstFactory; }
  IAddConstToVectorLengthFactory* getAddConstToVectorLengthFactory() noexcept final { return mAddToMagnitudeFactory; }
  ICosineSourceFactory* getCosineSourceFactory() noexcept final { return mCosineFactory; }
  IFileReaderFactory* getFileReaderFactory() noexcept final { return mFileReaderFactory; }
  IFirFactory* getFirFactory() noexcept final { return mFirFactory; }
  IHackrfSourceFactory* getHackrfSourceFactory() noexcept final { return mHackrfFactory; }
  ICudaFilterFactory* getInt8ToFloatFactory() noexcept final { return mInt8Factory; }
  ICudaFilterFactory* getMagnitudeFactory() noexcept final { return mMagnitudeFactory; }
  ICudaFilterFactory* getMultiplyFactory() noexcept final { return mMultiplyFactory; }
  IQuadDemodFactory* getQuadDemodFactory() noexcept final { return mQuadDemodFactory; }
  IMemSet* getSysMemSet() noexcept final { return mSysMemSet; }
  ICudaMemSetFactory* getCudaMemSetFactory() noexcept final { return mCudaMemSetFactory; }
  ISteppingDriverFactory* getSteppingDriverFactory() noexcept final { return mSteppingFactory; }
  IFilterDriverFactory* getFilterDriverFactory() noexcept final { return mFilterDriverFactory; }
  IPortRemappingSinkFactory* getPortRemappingSinkFactory() noexcept final { return mRemapSinkFactory; }
  IPortRemappingSourceFactory* getPortRemappingSourceFactory() noexcept final { return mRemapSourceFactory; }
  IRfToPcmAudioFactory* getRfToPcmAudioFactory() noexcept final { return mRfToPcmFactory; }
  IReadByteCountMonitorFactory* getReadByteCountMonitorFactory() noexcept final { return mMonitorFactory; }
  IDriverToDiagramFactory* getDriverToDotFactory() noexcept final { return mDotFactory; }
  IBufferRangeFactory* getBufferRangeFactory() noexcept final { return mRangeFactory; }
  ICommandQueueFactory* getCommandQueueFactory() noexcept final { return mQueueFactory; }
  ICudaCommandQueueFactory* getCudaCommandQueueFactory() noexcept final { return mCudaQueueFactory; }

  Result<IBufferFactory> createBufferFactory(IAllocator* allocator) noexcept final {
    NON_NULL_PARAM_OR_RET(allocator);
    return makeRefResultNonNull<IBufferFactory>(newBufferFactory(allocator));
  }
  Result<IRelocatableResizableBufferFactory> createRelocatableResizableBufferFactory(IAllocator* allocator,
                                                                                     const IBufferCopier* copier) noexcept final {
    NON_NULL_PARAM_OR_RET(allocator);
    NON_NULL_PARAM_OR_RET(copier);
    return makeRefResultNonNull<IRelocatableResizableBufferFactory>(newRelocatableBufferFactory(allocator, copier));
  }
  Result<IBufferPool> createBufferPool(size_t maxBufferCount, size_t bufferSize, IBufferFactory* bufferFactory) noexcept final {
    NON_NULL_PARAM_OR_RET(bufferFactory);
    return makeRefResultNonNull<IBufferPool>(newBufferPool(maxBufferCount, bufferSize, bufferFactory));
  }
  Result<IBufferPoolFactory> createBufferPoolFactory(size_t maxBufferCount, IBufferFactory* bufferFactory) noexcept final {
    NON_NULL_PARAM_OR_RET(bufferFactory);
    return makeRefResultNonNull<IBufferPoolFactory>(newBufferPoolFactory(maxBufferCount, bufferFactory));
  }

  // the singleton lives for the whole process (reference Factories.cpp:187-191)
  void ref() const noexcept final {}
  void unref() const noexcept final {}

 private:
  ~Factories() final = default;
  ConstRef<IAllocator> mSysMemAllocator;
  ConstRef<IBufferCopier> mSysMemCopier;
  ConstRef<IMemSet> mSysMemSet;
  ConstRef<ICudaAllocatorFactory> mCudaAllocatorFactory;
  ConstRef<ICudaBufferCopierFactory> mCudaCopierFactory;
  ConstRef<ICudaMemSetFactory> mCudaMemSetFactory;
  ConstRef<IBufferRangeFactory> mRangeFactory;
  ConstRef<IBufferSliceFactory> mSliceFactory;
  ConstRef<IBufferUtil> mBufferUtil;
  ConstRef<IResizableBufferFactory> mResizableFactory;
  ConstRef<ICudaCommandQueueFactory> mCudaQueueFactory;
  ConstRef<ICommandQueueFactory> mQueueFactory;
  ConstRef<ICudaMemcpyFilterFactory> mMemcpyFactory;
  ConstRef<IAacFileWriterFactory> mAacFactory;
  ConstRef<IAddConstFactory> mAddConstFactory;
  ConstRef<IAddConstToVectorLengthFactory> mAddToMagnitudeFactory;
  ConstRef<ICosineSourceFactory> mCosineFactory;
  ConstRef<IFileReaderFactory> mFileReaderFactory;
  ConstRef<IFirFactory> mFirFactory;
  ConstRef<IHackrfSourceFactory> mHackrfFactory;
  ConstRef<ICudaFilterFactory> mInt8Factory;
  ConstRef<ICudaFilterFactory> mMagnitudeFactory;
  ConstRef<ICudaFilterFactory> mMultiplyFactory;
  ConstRef<IQuadDemodFactory> mQuadDemodFactory;
  ConstRef<ISteppingDriverFactory> mSteppingFactory;
  ConstRef<IFilterDriverFactory> mFilterDriverFactory;
  ConstRef<IPortRemappingSinkFactory> mRemapSinkFactory;
  ConstRef<IPortRemappingSourceFactory> mRemapSourceFactory;
  ConstRef<IRfToPcmAudioFactory> mRfToPcmFactory;
  ConstRef<IReadByteCountMonitorFactory> mMonitorFactory;
  ConstRef<IDriverToDiagramFactory> mDotFactory;
};

std::mutex g_registryMutex;
std::map<std::string, Ref<INodeFactory>>& registry() {
  static std::map<std::string, Ref<INodeFactory>> r;
  return r;
}
std::once_flag g_defaultsOnce;

}  // namespace

GS_EXPORT Result<IFactories> getFactoriesSingleton() noexcept {
  static IFactories* const instance = new (std::nothrow) Factories();
  return makeRefResultNonNull<IFactories>(instance);
}

GS_EXPORT Status registerNodeFactory(const char* name, INodeFactory* factory) noexcept {
  GS_REQUIRE_OR_RET_STATUS(name != nullptr && factory != nullptr, "a name and a factory are required");
  try {
    std::lock_guard<std::mutex> lock(g_registryMutex);
    registry()[name] = factory;
    return Status_Success;
  }
  IF_CATCH_RETURN_STATUS
}

// names of the reference's registry (FilterFactories.cpp:132-150)
GS_EXPORT Status registerDefaultNodeFactories() noexcept {
  Result<IFactories> r = getFactoriesSingleton();
  if (r.status != Status_Success) return r.status;
  IFactories* f = r.value;
  Status result = Status_Success;
  std::call_once(g_defaultsOnce, [&]() {
    const std::pair<const char*, INodeFactory*> defaults[] = {
        {"AacFileWriter", f->getAacFileWriterFactory()},
        {"AacWriter", f->getAacFileWriterFactory()},      // FilterFactories.cpp:136
        {"Cosine", f->getCosineSourceFactory()},          // :140
        {"File", f->getFileReaderFactory()},              // :141
        {"HackRfSource", f->getHackrfSourceFactory()},    // :143
        {"AddConst", f->getAddConstFactory()},
        {"AddConstToVectorLength", f->getAddConstToVectorLengthFactory()},
        {"CosineSource", f->getCosineSourceFactory()},
        {"CudaMemcpy", f->getCudaMemcpyFilterFactory()},
        {"FileReader", f->getFileReaderFactory()},
        {"Fir", f->getFirFactory()},
        {"HackrfSource", f->getHackrfSourceFactory()},
        {"Int8ToFloat", f->getInt8ToFloatFactory()},
        {"Magnitude", f->getMagnitudeFactory()},
        {"MultiplyCCC", f->getMultiplyFactory()},
        {"Multiply", f->getMultiplyFactory()},  // the name RfToPcmAudioFactory.cpp:239-246 emits
        {"QuadDemod", f->getQuadDemodFactory()},
        {"RfToPcmAudio", f->getRfToPcmAudioFactory()},
        {"Component", f->getFilterDriverFactory()},
    };
    for (const auto& d : defaults) {
      const Status st = registerNodeFactory(d.first, d.second);
      if (st != Status_Success) result = st;
    }
  });
  return result;
}
GS_EXPORT Status registerDefaultFilterFactories() noexcept { return registerDefaultNodeFactories(); }

GS_EXPORT bool hasNodeFactory(const char* name) noexcept {
  if (name == nullptr || registerDefaultNodeFactories() != Status_Success) return false;
  std::lock_guard<std::mutex> lock(g_registryMutex);
  return registry().find(name) != registry().end();
}

GS_EXPORT Result<Node> createNode(const char* name, const char* jsonParameters) noexcept {
  NON_NULL_PARAM_OR_RET(name);
  FWD_IN_RESULT_IF_ERR(registerDefaultNodeFactories());
  Ref<INodeFactory> factory;
  {
    std::lock_guard<std::mutex> lock(g_registryMutex);
    const auto it = registry().find(name);
    if (it == registry().end()) {
      gsloge("No node factory is registered under [%s]", name);
      return ERR_RESULT(Status_NotFound);
    }
    factory = it->second;
  }
  return factory.get()->create(jsonParameters);
}

#define GS_CREATE_AS(Type__, cast__)                                                         \
  Result<Node> made = createNode(name, jsonParameters);                                      \
  if (made.status != Status_Success) return ERR_RESULT(made.status);                         \
  Type__* typed = made.value->cast__();                                                      \
  if (typed == nullptr) {                                                                    \
    gsloge("Node [%s] is not a " #Type__, name);                                             \
    made.value->unref();                                                                     \
    return ERR_RESULT(Status_InvalidArgument);                                               \
  }                                                                                          \
  return makeRefResultNonNull<Type__>(typed)

GS_EXPORT Result<Filter> createFilter(const char* name, const char* jsonParameters) noexcept { GS_CREATE_AS(Filter, asFilter); }
GS_EXPORT Result<Source> createSource(const char* name, const char* jsonParameters) noexcept { GS_CREATE_AS(Source, asSource); }
GS_EXPORT Result<Sink> createSink(const char* name, const char* jsonParameters) noexcept { GS_CREATE_AS(Sink, asSink); }
