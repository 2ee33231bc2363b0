// Command queues: one blocking CUDA stream per queue on a given device, plus the string-keyed registry that the JSON
// factories use to find a queue by id.  Reference: src/commandqueue/CudaCommandQueueFactory.h:14-20,
// src/commandqueue/CommandQueueFactory.cpp:35-77.
#include <map>
#include <mutex>

#include "internal.h"
#include "json_min.h"

namespace gs {
namespace {

class CudaCommandQueue final : public ICudaCommandQueue {
 public:
  CudaCommandQueue(int32_t device, cudaStream_t stream) noexcept : mDevice(device), mStream(stream) {}
  int32_t cudaDevice() const noexcept final { return mDevice; }
  cudaStream_t cudaStream() const noexcept final { return mStream; }

 private:
  const int32_t mDevice;
  const cudaStream_t mStream;
  ~CudaCommandQueue() final {
    CudaDevicePushPop device(mDevice);
    SAFE_CUDA_WARN_ONLY(cudaStreamDestroy(mStream));
  }
  REF_COUNTED_NO_DESTRUCTOR(CudaCommandQueue);
};

class CudaQueueFactory final : public ICudaCommandQueueFactory {
 public:
  CudaQueueFactory() noexcept = default;
  Result<ICudaCommandQueue> create(int32_t device) noexcept final {
    CUDA_DEV_PUSH_POP_OR_RET_RESULT(device);
    cudaStream_t stream = nullptr;
    // a BLOCKING stream, like the reference: its tests synchronise through the legacy default stream
    SAFE_CUDA_OR_RET_RESULT(cudaStreamCreate(&stream));
    return makeRefResultNonNull<ICudaCommandQueue>(new (std::nothrow) CudaCommandQueue(device, stream));
  }
  REF_COUNTED(CudaQueueFactory);
};

class NamedQueues final : public ICommandQueueFactory {
 public:
  explicit NamedQueues(IFactories* factories) noexcept : mFactories(factories) {}

  Status create(const char* queueId, const char* parameterJson) noexcept final {
    GS_REQUIRE_OR_RET_STATUS(queueId != nullptr && parameterJson != nullptr, "queue id and parameters are required");
    try {
      const Json params = Json::parse(parameterJson);
      const std::string type = params.at("queueType").str();
      GS_REQUIRE_OR_RET_STATUS_FMT(type == "cuda", "Unknown queue type [%s]", type.c_str());
      const int32_t device = static_cast<int32_t>(params.at("cudaDevice").num());
      std::lock_guard<std::mutex> lock(mMutex);
      GS_REQUIRE_OR_RET_STATUS_FMT(mQueues.find(queueId) == mQueues.end(), "Queue [%s] already exists", queueId);
      Ref<ICudaCommandQueue> queue;
      UNWRAP_OR_FWD_STATUS(queue, mFactories->getCudaCommandQueueFactory()->create(device));
      mQueues.emplace(queueId, queue);
      return Status_Success;
    } catch (const std::invalid_argument& e) {
      gsloge("Cannot create queue [%s]: %s", queueId, e.what());
      return Status_ParseError;
    }
    IF_CATCH_RETURN_STATUS
  }
  bool exists(const char* queueId) noexcept final {
    std::lock_guard<std::mutex> lock(mMutex);
    return queueId != nullptr && mQueues.find(queueId) != mQueues.end();
  }
  Result<ICudaCommandQueue> getCudaCommandQueue(const char* queueId) noexcept final {
    std::lock_guard<std::mutex> lock(mMutex);
    const auto it = queueId == nullptr ? mQueues.end() : mQueues.find(queueId);
    if (it == mQueues.end()) {
      gsloge("Command queue [%s] does not exist", queueId ? queueId : "(null)");
      return ERR_RESULT(Status_NotFound);
    }
    return makeRefResultNonNull<ICudaCommandQueue>(it->second.get());
  }

 private:
  IFactories* const mFactories;  // the singleton outlives everything
  std::mutex mMutex;
  std::map<std::string, Ref<ICudaCommandQueue>> mQueues;
  REF_COUNTED(NamedQueues);
};

}  // namespace

ICudaCommandQueueFactory* newCudaCommandQueueFactory() noexcept { return new (std::nothrow) CudaQueueFactory(); }
ICommandQueueFactory* newCommandQueueFactory(IFactories* factories) noexcept { return new (std::nothrow) NamedQueues(factories); }

}  // namespace gs
