// Stand-ins for the three factories of the reference whose implementations need libraries that are
// absent from this image (libhackrf, FFmpeg, kernrj/remez-exchange).  build_ref.sh substitutes this
// header for their #include lines in a scratch copy of /root/reference/src/Factories.cpp, so that the
// rest of the reference's host framework builds unmodified.  TEST INFRASTRUCTURE (oracle/), not product.
#pragma once

#include <ParseJson.h>  // the headers this file stands in for pull these in for the factories included after them
#include <gpusdrpipeline/Factories.h>
#include <gpusdrpipeline/filters/FilterFactories.h>

#include <nlohmann/json.hpp>

class AacFileWriterFactory final : public IAacFileWriterFactory {
 public:
  explicit AacFileWriterFactory(IFactories*) noexcept {}
  Result<Node> create(const char*) noexcept final { return {Status_NotFound, nullptr}; }
  Result<Sink> createAacFileWriter(const char*, int32_t, int32_t, ICudaCommandQueue*) noexcept final {
    return {Status_NotFound, nullptr};
  }
  REF_COUNTED(AacFileWriterFactory);
};

class HackrfSourceFactory final : public IHackrfSourceFactory {
 public:
  explicit HackrfSourceFactory(IFactories*) noexcept {}
  Result<Node> create(const char*) noexcept final { return {Status_NotFound, nullptr}; }
  Result<IHackrfSource> createHackrfSource(int32_t, uint64_t, double, size_t) noexcept final {
    return {Status_NotFound, nullptr};
  }
  REF_COUNTED(HackrfSourceFactory);
};

class RfToPcmAudioFactory final : public IRfToPcmAudioFactory {
 public:
  explicit RfToPcmAudioFactory(IFactories*) noexcept {}
  Result<Node> create(const char*) noexcept final { return {Status_NotFound, nullptr}; }
  Result<Filter> createRfToPcm(float, Modulation, size_t, size_t, float, float, float, float, float, float,
                               const char*) noexcept final {
    return {Status_NotFound, nullptr};
  }
  REF_COUNTED(RfToPcmAudioFactory);
};
