// ChainFilter -- the whole int8/cf32 -> mix -> FIR -> demod -> audio FIR path as ONE Filter node, running the fused
// sm_100a kernels behind include/b200sdr/b200sdr.h.  It is what IRfToPcmAudioFactory::createRfToPcm() returns in this
// library (the reference assembles a five-node Component: factories/RfToPcmAudioFactory.cpp:214-304) and what
// gsCreateFusedChain() (include/gpusdrpipeline/FusedChain.h) returns for caller-supplied taps.
//
// Stream semantics are those of the cascade it replaces: with N samples committed so far the node has produced
//   floor((floor((N - T1 + 1) / D1) - fm - T2 + 1) / D2)      outputs        (Fir.cpp:141-187, QuadFmDemod.cpp:76-84)
// because every readOutput consumes exactly nOut*D1*D2 input samples and leaves the rest -- the (T1-1) + (T2-1+fm)*D1
// samples of history plus any tail -- in the port buffer.  The mixer phase is a function of the ABSOLUTE sample index
// (64-bit fixed-point turns), so it neither drifts nor depends on how the stream is chunked.
#include <b200sdr/b200sdr.h>
#include <gpusdrpipeline/EventPipeline.h>
#include <gpusdrpipeline/FusedChain.h>

#include <algorithm>

#include <cmath>
#include <memory>
#include <vector>

#include "internal.h"
#include "json_min.h"
#include "port_input.h"

namespace gs {
namespace {

Status statusOf(b200sdr_status st) noexcept {
  if (st != B200SDR_OK) gsloge("b200sdr: %s", b200sdr_last_error());
  return static_cast<Status>(st);  // b200sdr status codes are the reference's Status values
}

class ChainFilter final : public Filter {
 public:
  static Result<Filter> create(const GsFusedChainParams& p, ICudaCommandQueue* queue, IFactories* f) noexcept {
    NON_NULL_PARAM_OR_RET(queue);
    GS_REQUIRE_OR_RET_RESULT(p.inputType == SampleType_Int8Complex || p.inputType == SampleType_FloatComplex,
                             "The fused chain takes int8-complex or float-complex input");
    GS_REQUIRE_OR_RET_RESULT(p.modulation == Modulation_Am || p.modulation == Modulation_Fm, "Unknown modulation");
    GS_REQUIRE_OR_RET_RESULT(p.rfTaps != nullptr && p.rfTapCount > 0 && p.audioTaps != nullptr && p.audioTapCount > 0,
                             "RF and audio taps are required");
    ChainFilter* node = new (std::nothrow) ChainFilter(queue, p.inputType == SampleType_Int8Complex ? 2 : 8);
    NON_NULL_OR_RET(node);
    Status st = node->init(p, f);
    if (st != Status_Success) {
      node->unref();
      return ERR_RESULT(st);
    }
    return makeRefResultNonNull<Filter>(node);
  }

  Result<IBuffer> requestBuffer(size_t port, size_t byteCount) noexcept final {
    GS_REQUIRE_OR_RET_RESULT_FMT(port == 0, "Cannot request buffer. Input port [%zu] is out of range.", port);
    return mPort->request(byteCount);
  }
  Status commitBuffer(size_t port, size_t byteCount) noexcept final {
    GS_REQUIRE_OR_RET_STATUS_FMT(port == 0, "Cannot commit buffer. Input port [%zu] is out of range", port);
    return mPort->commit(byteCount);
  }
  // large steps: the kernels are persistent and a step's cost is dominated by its launch below a few MiB of input
  size_t preferredInputBufferSize(size_t) noexcept final { return size_t(1) << 26; }
  size_t getOutputDataSize(size_t port) noexcept final { return port == 0 ? numOutputs() * sizeof(float) : 0; }
  size_t getOutputSizeAlignment(size_t port) noexcept final { return port == 0 ? 32 * sizeof(float) : 0; }
  IBufferCopier* getOutputCopier(size_t port) noexcept final { return port == 0 ? mD2d.get().get() : nullptr; }

  Status readOutput(IBuffer** bufs, size_t) noexcept final {
    GS_REQUIRE_OR_RET_STATUS(bufs != nullptr && bufs[0] != nullptr, "One output port is required");
    IBuffer* out = bufs[0];
    size_t n = numOutputs();
    const size_t room = out->range()->remaining() / sizeof(float);
    if (n > room) n = room;
    if (n == 0) return Status_Success;
    const size_t nIn = mPort->used() / mElemBytes;
    const size_t demodCount = (n - 1) * mD2 + mT2;
    if (demodCount > mScratchCount) {  // only the two-kernel fallback touches it
      UNWRAP_OR_FWD_STATUS(mScratch, mAllocator.get()->allocate(demodCount * sizeof(float)));
      mScratchCount = demodCount;
    }
    FWD_IF_ERR(statusOf(b200sdr_chain_run(mChain, mPort->data(), nIn, mAbsoluteIndex, mScratch.get()->as<float>(), out->writePtr<float>(), n,
                                          mQueue->cudaStream())));
    FWD_IF_ERR(out->range()->increaseEndOffset(n * sizeof(float)));
    const size_t consumed = n * mStride;
    mPort->consume(consumed * mElemBytes);
    mAbsoluteIndex += consumed;
    return Status_Success;
  }

 private:
  ChainFilter(ICudaCommandQueue* queue, size_t elemBytes) noexcept : mQueue(queue), mElemBytes(elemBytes) {}
  ~ChainFilter() final {
    if (mChain != nullptr) b200sdr_chain_destroy(mChain);
  }
  Status init(const GsFusedChainParams& p, IFactories* f) noexcept {
    UNWRAP_OR_FWD_STATUS(mAllocator, f->getCudaAllocatorFactory()->createCudaAllocator(mQueue, 256, false));
    UNWRAP_OR_FWD_STATUS(mD2d, f->getCudaBufferCopierFactory()->createBufferCopier(mQueue, cudaMemcpyDeviceToDevice));
    mPort.reset(new (std::nothrow) PortInput(mAllocator.get(), mD2d.get(), mQueue, false));
    GS_REQUIRE_OR_RET(mPort != nullptr, "out of memory", Status_OutOfMemory);
    b200sdr_chain_config cfg {};
    cfg.struct_size = sizeof(cfg);
    cfg.input_type = p.inputType == SampleType_Int8Complex ? B200SDR_INPUT_INT8 : B200SDR_INPUT_CF32;
    cfg.modulation = p.modulation == Modulation_Fm ? B200SDR_MOD_FM : B200SDR_MOD_AM;
    cfg.mix = p.mix;
    cfg.sample_rate = p.sampleRate;
    cfg.frequency = p.frequency;
    cfg.rf_taps = p.rfTaps;
    cfg.rf_tap_count = p.rfTapCount;
    cfg.rf_decimation = p.rfDecimation;
    cfg.fm_gain = p.fmGain;
    cfg.audio_taps = p.audioTaps;
    cfg.audio_tap_count = p.audioTapCount;
    cfg.audio_decimation = p.audioDecimation;
    cfg.cuda_device = mQueue->cudaDevice();
    FWD_IF_ERR(statusOf(b200sdr_chain_create(&cfg, &mChain)));
    mStride = b200sdr_chain_input_stride(mChain);
    mT2 = p.audioTapCount;
    mD2 = p.audioDecimation == 0 ? 1 : p.audioDecimation;
    gslogd("Fused chain node: %s", b200sdr_chain_variant(mChain));
    return Status_Success;
  }
  size_t numOutputs() const noexcept {
    size_t audio = 0;
    b200sdr_chain_counts(mChain, mPort->used() / mElemBytes, nullptr, nullptr, &audio);
    return audio;
  }

  ConstRef<ICudaCommandQueue> mQueue;
  const size_t mElemBytes;
  Ref<IAllocator> mAllocator;
  Ref<IBufferCopier> mD2d;
  std::unique_ptr<PortInput> mPort;
  b200sdr_chain* mChain = nullptr;
  Ref<IMemory> mScratch;
  size_t mScratchCount = 0, mStride = 1, mT2 = 1, mD2 = 1;
  uint64_t mAbsoluteIndex = 0;
  REF_COUNTED_NO_DESTRUCTOR(ChainFilter);
};

// ---- tap design for createRfToPcm ------------------------------------------------------------------------------
// The reference designs equiripple taps with the un-vendored kernrj/remez-exchange (RfToPcmAudioFactory.cpp:49-122).
// This library uses a Kaiser-windowed sinc of the reference's own length estimate (fred harris,
// RfToPcmAudioFactory.cpp:44-47) and the requested stop-band attenuation; unity DC gain; fp64 -> fp32.
double besselI0(double x) {
  double sum = 1.0, term = 1.0;
  for (int k = 1; k < 64; k++) {
    term *= (x / (2.0 * k)) * (x / (2.0 * k));
    sum += term;
    if (term < 1e-18 * sum) break;
  }
  return sum;
}

std::vector<float> kaiserLowPass(double sampleRate, double cutoff, double transition, double dbAttenuation) {
  const double atten = std::fabs(dbAttenuation);
  size_t count = static_cast<size_t>(std::lrint(std::ceil(atten / (22.0 * (transition / sampleRate)))));
  if (count < 3) count = 3;
  if (count % 2 == 0) count++;  // type-I linear phase
  const double beta = atten > 50.0 ? 0.1102 * (atten - 8.7) : atten >= 21.0 ? 0.5842 * std::pow(atten - 21.0, 0.4) + 0.07886 * (atten - 21.0) : 0.0;
  const double fc = (cutoff + 0.5 * transition) / sampleRate;  // -6 dB point in the middle of the transition band
  const double mid = 0.5 * static_cast<double>(count - 1);
  std::vector<double> h(count);
  double sum = 0.0;
  for (size_t i = 0; i < count; i++) {
    const double t = static_cast<double>(i) - mid;
    const double sinc = t == 0.0 ? 2.0 * fc : std::sin(2.0 * M_PI * fc * t) / (M_PI * t);
    const double r = mid == 0.0 ? 0.0 : t / mid;
    h[i] = sinc * besselI0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / besselI0(beta);
    sum += h[i];
  }
  std::vector<float> out(count);
  for (size_t i = 0; i < count; i++) out[i] = static_cast<float>(h[i] / sum);
  return out;
}

// Parameter derivation follows RfToPcmAudioFactory.cpp:164-170 (cut-offs at 95 % / 90 % of the output Nyquist
// frequencies, 5 % / 10 % transitions), except that the audio low-pass is designed at the rate it runs at (the
// demodulator's output rate) -- the reference designs it at the audio rate, :185-190, which makes it an all-pass.
void designRfToPcmTaps(float rfSampleRate, size_t rfDecim, size_t audioDecim, float rfDbAttenuation, float audioDbAttenuation,
                       std::vector<float>& rfTaps, std::vector<float>& audioTaps) {
  const double demodRate = static_cast<double>(rfSampleRate) / static_cast<double>(rfDecim);
  const double audioRate = demodRate / static_cast<double>(audioDecim);
  rfTaps = kaiserLowPass(rfSampleRate, demodRate / 2.0 * 0.95, demodRate / 2.0 * 0.05, rfDbAttenuation);
  audioTaps = kaiserLowPass(demodRate, audioRate / 2.0 * 0.9, audioRate / 2.0 * 0.1, audioDbAttenuation);
}

class RfToPcmFactory final : public IRfToPcmAudioFactory {
 public:
  explicit RfToPcmFactory(IFactories* f) noexcept : mFactories(f) {}

  Result<Filter> createRfToPcm(float rfSampleRate, Modulation modulation, size_t rfDecim, size_t audioDecim, float centerFrequency,
                               float channelFrequency, float channelWidth, float fskDeviationIfFm, float rfDbAttenuation,
                               float audioDbAttenuation, const char* commandQueueId) noexcept final {
    return build(SampleType_FloatComplex, rfSampleRate, modulation, rfDecim, audioDecim, centerFrequency, channelFrequency, channelWidth,
                 fskDeviationIfFm, rfDbAttenuation, audioDbAttenuation, commandQueueId);
  }

  // JSON keys of the reference (RfToPcmAudioFactory.cpp:131-150) plus the additive "inputType": "Int8Complex"
  Result<Node> create(const char* json) noexcept final {
    try {
      const Json p = Json::parse(json);
      const std::string& mod = p.at("modulation").str();
      const bool fm = mod == "fm" || mod == "FM";
      GS_REQUIRE_OR_RET_RESULT_FMT(fm || mod == "am" || mod == "AM", "Unknown modulation [%s]", mod.c_str());
      SampleType inputType = SampleType_FloatComplex;
      if (p.contains("inputType") && (p.at("inputType").str() == "Int8Complex" || p.at("inputType").str() == "int8Complex"))
        inputType = SampleType_Int8Complex;
      auto num = [&](const char* key) { return static_cast<float>(p.at(key).num()); };
      return ResultCast<Node>(build(inputType, num("rfSampleRate"), fm ? Modulation_Fm : Modulation_Am,
                                    static_cast<size_t>(p.at("rfLowPassDecimation").num()),
                                    static_cast<size_t>(p.at("audioLowPassDecimation").num()), num("tunedFrequency"), num("channelFrequency"),
                                    num("channelWidth"), fm ? num("fskDeviation") : 0.0f, num("rfLowPassDbAttenuation"),
                                    num("audioLowPassDbAttenuation"), p.at("commandQueue").str().c_str()));
    } catch (const std::invalid_argument& e) {
      gsloge("Bad RfToPcmAudio parameters: %s", e.what());
      return ERR_RESULT(Status_ParseError);
    }
    IF_CATCH_RETURN_RESULT
  }

 private:
  Result<Filter> build(SampleType inputType, float rfSampleRate, Modulation modulation, size_t rfDecim, size_t audioDecim, float centerFrequency,
                       float channelFrequency, float channelWidth, float fskDeviationIfFm, float rfDbAttenuation, float audioDbAttenuation,
                       const char* commandQueueId) noexcept {
    try {
      GS_REQUIRE_OR_RET_RESULT(rfDecim > 0 && audioDecim > 0 && rfSampleRate > 0.0f, "rates and decimations must be positive");
      Ref<ICudaCommandQueue> queue;
      UNWRAP_OR_FWD_RESULT(queue, mFactories->getCommandQueueFactory()->getCudaCommandQueue(commandQueueId));
      const double demodRate = static_cast<double>(rfSampleRate) / static_cast<double>(rfDecim);
      std::vector<float> rfTaps, audioTaps;
      designRfToPcmTaps(rfSampleRate, rfDecim, audioDecim, rfDbAttenuation, audioDbAttenuation, rfTaps, audioTaps);
      GsFusedChainParams p {};
      p.structSize = sizeof(p);
      p.inputType = inputType;
      p.modulation = modulation;
      p.mix = 1;
      p.sampleRate = rfSampleRate;
      p.frequency = static_cast<double>(centerFrequency) - static_cast<double>(channelFrequency);  // RfToPcmAudioFactory.cpp:225
      p.rfTaps = rfTaps.data();
      p.rfTapCount = rfTaps.size();
      p.rfDecimation = rfDecim;
      // the Component passes the channel width as the deviation of its QuadDemod node (RfToPcmAudioFactory.cpp:262-270)
      const float deviation = fskDeviationIfFm > 0.0f ? fskDeviationIfFm : channelWidth;
      p.fmGain = static_cast<float>(demodRate) / (2.0f * static_cast<float>(M_PI) * deviation * 5);  // QuadDemodFactory.h:108-110
      p.audioTaps = audioTaps.data();
      p.audioTapCount = audioTaps.size();
      p.audioDecimation = audioDecim;
      gslogd("RfToPcm: %zu RF taps / %zu, %zu audio taps / %zu, shift %.1f Hz", rfTaps.size(), rfDecim, audioTaps.size(), audioDecim, p.frequency);
      return ChainFilter::create(p, queue.get(), mFactories);
    }
    IF_CATCH_RETURN_RESULT
  }
  IFactories* const mFactories;
  REF_COUNTED(RfToPcmFactory);
};

}  // namespace

IRfToPcmAudioFactory* newRfToPcmAudioFactory(IFactories* f) noexcept { return new (std::nothrow) RfToPcmFactory(f); }

Status designRfToPcmTapsC(float rfSampleRate, size_t rfDecim, size_t audioDecim, float rfDb, float audioDb, float* rfTaps, size_t rfCapacity,
                          size_t* rfCount, float* audioTaps, size_t audioCapacity, size_t* audioCount) noexcept {
  try {
    GS_REQUIRE_OR_RET_STATUS(rfDecim > 0 && audioDecim > 0 && rfSampleRate > 0.0f, "rates and decimations must be positive");
    std::vector<float> rf, audio;
    designRfToPcmTaps(rfSampleRate, rfDecim, audioDecim, rfDb, audioDb, rf, audio);
    if (rfCount != nullptr) *rfCount = rf.size();
    if (audioCount != nullptr) *audioCount = audio.size();
    if ((rfTaps != nullptr && rfCapacity < rf.size()) || (audioTaps != nullptr && audioCapacity < audio.size())) return Status_OutOfRange;
    if (rfTaps != nullptr) std::copy(rf.begin(), rf.end(), rfTaps);
    if (audioTaps != nullptr) std::copy(audio.begin(), audio.end(), audioTaps);
    return Status_Success;
  }
  IF_CATCH_RETURN_STATUS
}

// ---- IEventPipeline (include/gpusdrpipeline/EventPipeline.h; reference Waiter.cpp:34-50) ---------------------------
class EventPipeline final : public IEventPipeline {
 public:
  explicit EventPipeline(ICudaCommandQueue* queue) noexcept : mQueue(queue) {}
  Status recordNextAndWaitPrevious() noexcept final {
    CUDA_DEV_PUSH_POP_OR_RET_STATUS(mQueue->cudaDevice());
    if (mNext == nullptr) SAFE_CUDA_OR_RET_STATUS(cudaEventCreateWithFlags(&mNext, cudaEventDisableTiming));
    SAFE_CUDA_OR_RET_STATUS(cudaEventRecord(mNext, mQueue->cudaStream()));
    if (mPending) SAFE_CUDA_OR_RET_STATUS(cudaEventSynchronize(mPrevious));
    std::swap(mPrevious, mNext);
    mPending = true;
    return Status_Success;
  }
  Status waitLast() noexcept final {
    if (!mPending) return Status_Success;
    CUDA_DEV_PUSH_POP_OR_RET_STATUS(mQueue->cudaDevice());
    SAFE_CUDA_OR_RET_STATUS(cudaEventSynchronize(mPrevious));
    mPending = false;
    return Status_Success;
  }

 private:
  ~EventPipeline() final {
    if (mPrevious != nullptr || mNext != nullptr) {
      CudaDevicePushPop device(mQueue->cudaDevice());
      if (mPrevious != nullptr) cudaEventDestroy(mPrevious);
      if (mNext != nullptr) cudaEventDestroy(mNext);
    }
  }
  ConstRef<ICudaCommandQueue> mQueue;
  cudaEvent_t mPrevious = nullptr, mNext = nullptr;
  bool mPending = false;
  REF_COUNTED_NO_DESTRUCTOR(EventPipeline);
};

}  // namespace gs

GS_EXPORT Status gsDesignRfToPcmTaps(float rfSampleRate, size_t rfDecim, size_t audioDecim, float rfDb, float audioDb, float* rfTaps,
                                     size_t rfCapacity, size_t* rfCount, float* audioTaps, size_t audioCapacity, size_t* audioCount) noexcept {
  return gs::designRfToPcmTapsC(rfSampleRate, rfDecim, audioDecim, rfDb, audioDb, rfTaps, rfCapacity, rfCount, audioTaps, audioCapacity, audioCount);
}

GS_EXPORT Result<IEventPipeline> gsCreateEventPipeline(ICudaCommandQueue* commandQueue) noexcept {
  NON_NULL_PARAM_OR_RET(commandQueue);
  return makeRefResultNonNull<IEventPipeline>(new (std::nothrow) gs::EventPipeline(commandQueue));
}

GS_EXPORT Result<Filter> gsCreateFusedChain(const GsFusedChainParams* params, ICudaCommandQueue* commandQueue) noexcept {
  NON_NULL_PARAM_OR_RET(params);
  GS_REQUIRE_OR_RET_RESULT(params->structSize == sizeof(GsFusedChainParams), "GsFusedChainParams::structSize mismatch");
  Result<IFactories> factories = getFactoriesSingleton();
  if (factories.status != Status_Success) return ERR_RESULT(factories.status);
  return gs::ChainFilter::create(*params, commandQueue, factories.value);
}
