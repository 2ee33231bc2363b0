// The fused chain object behind include/b200sdr/b200sdr.h:
//   K1 (rowsKernel or directKernel): int8/cf32 -> [mix] -> decimating FIR -> AM/FM demod, one pass over HBM
//   K2 (directKernel, staged):       audio-rate FIR (float taps, float data, decimating)
// plus the host-buffer path (overlapped time segments staged through double-buffered device memory).
#include <b200sdr/b200sdr.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "chain_dispatch.h"
#include "digits.h"
#include "chain_kernels.cuh"
#include "fir_dispatch.h"
#include "fir_kernels.cuh"
#include "toeplitz_dispatch.h"
#include "toeplitz_kernels.cuh"

using namespace b200sdr;

namespace {

thread_local std::string t_lastError;

b200sdr_status fail(b200sdr_status status, const std::string& what) {
  t_lastError = what;
  return status;
}

b200sdr_status cudaFail(cudaError_t e, const char* where) {
  t_lastError = std::string(where) + ": " + cudaGetErrorString(e);
  return e == cudaErrorMemoryAllocation ? B200SDR_OUT_OF_MEMORY : B200SDR_RUNTIME_ERROR;
}

}  // namespace

namespace b200sdr {
b200sdr_status chainFail(b200sdr_status status, const std::string& what) { return fail(status, what); }
}  // namespace b200sdr

namespace {

#define CUDA_OR_RETURN(call)                              \
  do {                                                    \
    const cudaError_t e__ = (call);                       \
    if (e__ != cudaSuccess) return cudaFail(e__, #call);  \
  } while (false)

size_t firNumOutputs(size_t nIn, size_t T, size_t D) {
  if (D == 0) D = 1;
  if (T == 0 || nIn + 1 < T) return 0;
  return (nIn + 1 - T) / D;
}

float2 hostPhasor(uint64_t turns) {
  const double frac = static_cast<double>(static_cast<int64_t>(turns)) * (1.0 / 18446744073709551616.0);
  const double phi = 6.283185307179586476925286766559 * frac;
  return make_float2(static_cast<float>(std::cos(phi)), static_cast<float>(std::sin(phi)));
}

constexpr int kSlots = 2;

}  // namespace

struct b200sdr_chain {
  int device = 0;
  int elem = kElemInt8Complex;
  int mod = kModAm;
  bool mix = true;
  unsigned T1 = 0, D1 = 1, T2 = 0, D2 = 1;
  uint64_t phaseStep = 0;
  float fmGain = 1.0f;
  float inScale = 1.0f;

  // device-resident constants
  float* dTaps1 = nullptr;
  float* dTaps2 = nullptr;
  float* dTapTable = nullptr;   // hT[D1][TS]
  float2* dMixTable = nullptr;  // W[D1]
  float2* dRotTable = nullptr;  // exp(j w m D1), m <= TS
  unsigned* dBFrag = nullptr;   // tensor route: int8 digit fragments of the mixer-rotated tap matrix
  float digitScale[3] = {0.0f, 0.0f, 0.0f};
  FirRoute tableRoute {};       // route the tables were laid out for (aligned input)
  ChainPlan fusedPlan {};       // fused persistent kernel (AM/FM with an audio FIR, aligned input), if the shape allows
  ToepPlan toepPlan {};         // int8 input: fused kernel whose RF stage is one int8 GEMM over a Toeplitz view of the input
  uint4* dToepFrag = nullptr;
  float* dToepTaps2 = nullptr;  // audio taps zero-padded to a multiple of 4 (one bulk copy in the kernel)
  float toepScale[3] = {0.0f, 0.0f, 0.0f};
  float2 rot1 = make_float2(1.0f, 0.0f);
  std::string variant;

  // host-buffer path
  size_t hostSegment = static_cast<size_t>(1) << 25;
  cudaStream_t h2d = nullptr, compute = nullptr, d2h = nullptr;
  cudaEvent_t copied[kSlots] {}, computed[kSlots] {}, drained[kSlots] {};
  void* slotIn[kSlots] {};
  float* slotDemod[kSlots] {};
  float* slotAudio[kSlots] {};
  size_t slotInBytes = 0, slotDemodCount = 0, slotAudioCount = 0;

  size_t elemBytes() const { return elem == kElemInt8Complex ? 2 : 8; }
  unsigned fm() const { return mod == kModFm ? 1u : 0u; }
  bool hasAudioFir() const { return T2 > 0 && mod != kModNone; }
  size_t outElemFloats() const { return mod == kModNone ? 2 : 1; }
  size_t stride() const { return static_cast<size_t>(D1) * (hasAudioFir() ? D2 : 1); }
  size_t window() const { return (static_cast<size_t>(hasAudioFir() ? T2 - 1 : 0) + fm()) * D1 + T1; }
};

// ---------------------------------------------------------------------------------------------------
B200SDR_EXPORT const char* b200sdr_last_error(void) { return t_lastError.c_str(); }
B200SDR_EXPORT const char* b200sdr_version(void) { return "b200sdr 0.1 (sm_100a)"; }
B200SDR_EXPORT uint64_t b200sdr_launch_count(void) { return g_launchCount.load(); }
B200SDR_EXPORT uint64_t b200sdr_phase_step(double frequency, double sampleRate) { return phaseStepOf(frequency, sampleRate); }
B200SDR_EXPORT size_t b200sdr_fir_num_outputs(size_t numInputs, size_t tapCount, size_t decimation) {
  return firNumOutputs(numInputs, tapCount, decimation);
}

B200SDR_EXPORT b200sdr_status b200sdr_toeplitz_tables(
    const float* rfTaps, size_t rfTapCount, size_t rfDecimation, uint32_t mix, double frequency, double sampleRate, uint32_t* fragments,
    size_t fragCapacityWords, size_t* fragWords, float digitScale[3], uint32_t* kSteps, uint32_t* magic) {
  if (!rfTaps || rfTapCount == 0 || rfTapCount > (1u << 24) || rfDecimation == 0 || rfDecimation % 8 != 0 || rfDecimation > (1u << 16))
    return fail(B200SDR_INVALID_ARGUMENT, "rfTaps must be non-null and rfDecimation a multiple of 8");
  if (mix && !(sampleRate > 0.0)) return fail(B200SDR_INVALID_ARGUMENT, "sample_rate must be positive");
  ToepPlan plan {};
  const unsigned T1 = static_cast<unsigned>(rfTapCount), D1 = static_cast<unsigned>(rfDecimation);
  plan.KS = (2u * T1 + 6u * D1 + 31u) / 32u;
  plan.Q = (plan.KS + 1u) / 2u;
  const size_t words = static_cast<size_t>(plan.Q) * 32u * 12u;
  if (fragWords) *fragWords = words;
  if (kSteps) *kSteps = plan.KS;
  if (!fragments) return B200SDR_OK;
  if (fragCapacityWords < words) return fail(B200SDR_OUT_OF_RANGE, "fragments buffer too small");
  std::vector<uint32_t> frag;
  float scale[3];
  buildToeplitzFragments(rfTaps, T1, D1, mix != 0, mix ? phaseStepOf(frequency, sampleRate) : 0, 1.0 / 128.0, plan, frag, scale);
  std::memcpy(fragments, frag.data(), words * sizeof(uint32_t));
  if (digitScale) std::memcpy(digitScale, scale, sizeof(scale));
  if (magic) *magic = plan.magic ? 1u : 0u;
  return B200SDR_OK;
}

B200SDR_EXPORT void b200sdr_chain_destroy(b200sdr_chain* c) {
  if (!c) return;
  DeviceGuard guard(c->device);
  for (int s = 0; s < kSlots; s++) {
    if (c->slotIn[s]) cudaFree(c->slotIn[s]);
    if (c->slotDemod[s]) cudaFree(c->slotDemod[s]);
    if (c->slotAudio[s]) cudaFree(c->slotAudio[s]);
    if (c->copied[s]) cudaEventDestroy(c->copied[s]);
    if (c->computed[s]) cudaEventDestroy(c->computed[s]);
    if (c->drained[s]) cudaEventDestroy(c->drained[s]);
  }
  if (c->h2d) cudaStreamDestroy(c->h2d);
  if (c->compute) cudaStreamDestroy(c->compute);
  if (c->d2h) cudaStreamDestroy(c->d2h);
  cudaFree(c->dTaps1);
  cudaFree(c->dTaps2);
  cudaFree(c->dTapTable);
  cudaFree(c->dMixTable);
  cudaFree(c->dRotTable);
  cudaFree(c->dBFrag);
  cudaFree(c->dToepFrag);
  cudaFree(c->dToepTaps2);
  delete c;
}

B200SDR_EXPORT b200sdr_status b200sdr_chain_create(const b200sdr_chain_config* cfg, b200sdr_chain** chainOut) {
  if (!cfg || !chainOut) return fail(B200SDR_INVALID_ARGUMENT, "config and chainOut must be non-null");
  *chainOut = nullptr;
  if (cfg->struct_size != sizeof(b200sdr_chain_config)) return fail(B200SDR_INVALID_ARGUMENT, "struct_size mismatch");
  if (cfg->input_type != B200SDR_INPUT_CF32 && cfg->input_type != B200SDR_INPUT_INT8)
    return fail(B200SDR_INVALID_ARGUMENT, "input_type must be float-complex (0) or int8-complex (2)");
  if (cfg->modulation > B200SDR_MOD_NONE) return fail(B200SDR_INVALID_ARGUMENT, "unknown modulation");
  if (!cfg->rf_taps || cfg->rf_tap_count == 0 || cfg->rf_tap_count > (1u << 24))
    return fail(B200SDR_INVALID_ARGUMENT, "rf_taps must be non-null with 1..2^24 taps");
  if (cfg->rf_decimation > (1u << 24) || cfg->audio_decimation > (1u << 24) || cfg->audio_tap_count > (1u << 24))
    return fail(B200SDR_INVALID_ARGUMENT, "decimation / tap count out of range");
  if (cfg->audio_tap_count > 0 && !cfg->audio_taps) return fail(B200SDR_INVALID_ARGUMENT, "audio_taps is null");
  if (cfg->mix && !(cfg->sample_rate > 0.0)) return fail(B200SDR_INVALID_ARGUMENT, "sample_rate must be positive");

  b200sdr_chain* c = new (std::nothrow) b200sdr_chain();
  if (!c) return fail(B200SDR_OUT_OF_MEMORY, "host allocation failed");
  c->device = cfg->cuda_device;
  c->elem = cfg->input_type == B200SDR_INPUT_INT8 ? kElemInt8Complex : kElemComplex;
  c->mod = static_cast<int>(cfg->modulation);
  c->mix = cfg->mix != 0;
  c->T1 = static_cast<unsigned>(cfg->rf_tap_count);
  c->D1 = cfg->rf_decimation == 0 ? 1u : static_cast<unsigned>(cfg->rf_decimation);  // Fir.cpp:119
  c->T2 = cfg->audio_taps ? static_cast<unsigned>(cfg->audio_tap_count) : 0u;
  c->D2 = cfg->audio_decimation == 0 ? 1u : static_cast<unsigned>(cfg->audio_decimation);
  c->phaseStep = c->mix ? phaseStepOf(cfg->frequency, cfg->sample_rate) : 0;
  c->fmGain = cfg->fm_gain;
  c->inScale = c->elem == kElemInt8Complex ? 1.0f / 128.0f : 1.0f;

  DeviceGuard guard(c->device);
  b200sdr_status st = B200SDR_OK;
  auto upload = [&](const void* host, size_t bytes, void** dev) -> bool {
    cudaError_t e = cudaMalloc(dev, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      st = cudaFail(e, "uploading chain constants");
      return false;
    }
    return true;
  };
  bool ok = guard.status == cudaSuccess;
  if (!ok) st = cudaFail(guard.status, "cudaSetDevice");
  ok = ok && upload(cfg->rf_taps, sizeof(float) * c->T1, reinterpret_cast<void**>(&c->dTaps1));
  if (ok && c->T2) ok = upload(cfg->audio_taps, sizeof(float) * c->T2, reinterpret_cast<void**>(&c->dTaps2));

  // Tables of the rows kernel, evaluated in fp64 on the host (for a 16-byte-aligned input pointer).
  c->tableRoute = planFir(c->elem, false, nullptr, c->T1, c->D1, c->mod);
  if (ok && c->tableRoute.rows) {
    const unsigned TS = c->tableRoute.TS, M = c->tableRoute.M, D = c->D1;
    std::vector<float> hT(static_cast<size_t>(D) * TS, 0.0f);
    for (unsigned p = 0; p < D; p++)
      for (unsigned m = 0; m < M; m++) {
        const size_t j = static_cast<size_t>(m) * D + p;
        if (j < c->T1) hT[static_cast<size_t>(p) * TS + m] = cfg->rf_taps[j] * (c->mix ? 1.0f : c->inScale);
      }
    ok = upload(hT.data(), hT.size() * sizeof(float), reinterpret_cast<void**>(&c->dTapTable));
    if (ok && c->mix) {
      std::vector<float2> W(D), rot(TS + 1);
      for (unsigned p = 0; p < D; p++) {
        W[p] = hostPhasor(c->phaseStep * p);
        W[p].x *= c->inScale;
        W[p].y *= c->inScale;
      }
      for (unsigned m = 0; m <= TS; m++) rot[m] = hostPhasor(c->phaseStep * (static_cast<uint64_t>(m) * D));
      ok = upload(W.data(), W.size() * sizeof(float2), reinterpret_cast<void**>(&c->dMixTable)) &&
           upload(rot.data(), rot.size() * sizeof(float2), reinterpret_cast<void**>(&c->dRotTable));
    }
  }
  if (!ok) {
    b200sdr_chain_destroy(c);
    return st;
  }
  char buf[256];
  if (c->hasAudioFir() && c->tableRoute.rows) {
    c->fusedPlan = planChain(c->elem, c->mix, nullptr, c->T1, c->D1, c->mod, c->T2, c->D2, c->device);
  }
  if (c->fusedPlan.fused && c->fusedPlan.conv) {
    // B[k][n] of the tensor route (chain_kernels.cuh): k = 2p + (0: I, 1: Q), n = 2m + (0: re, 1: im),
    // g[p][m] = h[m*D1 + p] * exp(j*w*p) / 128 evaluated in fp64, quantised to 24-bit fixed point = 3 signed int8 digits.
    const unsigned D = c->D1, M = c->fusedPlan.M, KS = c->fusedPlan.kSteps, NT = (2u * c->fusedPlan.MP + 7u) / 8u;
    const unsigned K = KS * 32u, N = NT * 8u;
    std::vector<double> B(static_cast<size_t>(K) * N, 0.0);
    double bMax = 0.0;
    for (unsigned p = 0; p < D; p++) {
      double wr = static_cast<double>(c->inScale), wi = 0.0;
      if (c->mix) {
        const double frac = static_cast<double>(static_cast<int64_t>(c->phaseStep * p)) * (1.0 / 18446744073709551616.0);
        const double phi = 6.283185307179586476925286766559 * frac;
        wr = std::cos(phi) * c->inScale;
        wi = std::sin(phi) * c->inScale;
      }
      for (unsigned m = 0; m < M; m++) {
        const size_t j = static_cast<size_t>(m) * D + p;
        const double h = j < c->T1 ? static_cast<double>(cfg->rf_taps[j]) : 0.0;
        const double gr = h * wr, gi = h * wi;
        B[static_cast<size_t>(2 * p) * N + 2 * m] = gr;
        B[static_cast<size_t>(2 * p + 1) * N + 2 * m] = -gi;
        B[static_cast<size_t>(2 * p) * N + 2 * m + 1] = gi;
        B[static_cast<size_t>(2 * p + 1) * N + 2 * m + 1] = gr;
      }
    }
    for (double v : B) bMax = std::fmax(bMax, std::fabs(v));
    const FixedPoint24 fx = fixedPoint24For(bMax);
    for (int d = 0; d < 3; d++) c->digitScale[d] = fx.digitScale[d];
    std::vector<unsigned> frag(c->fusedPlan.bFragWords, 0u);
    for (unsigned n8 = 0; n8 < NT; n8++)
      for (unsigned ks = 0; ks < KS; ks++)
        for (unsigned half = 0; half < 2; half++)
          for (unsigned lane = 0; lane < 32; lane++) {
            const unsigned g = lane >> 2, t = lane & 3u;
            unsigned word[3] = {0, 0, 0};
            for (unsigned e = 0; e < 4; e++) {
              const unsigned k = ks * 32u + half * 16u + t * 4u + e, n = n8 * 8u + g;
              int dg[3];
              balancedDigits(B[static_cast<size_t>(k) * N + n], fx.scale, dg);
              for (int d = 0; d < 3; d++) word[d] |= (static_cast<unsigned>(dg[d]) & 0xffu) << (8u * e);
            }
            for (unsigned d = 0; d < 3; d++) frag[(((d * NT + n8) * KS + ks) * 2u + half) * 32u + lane] = word[d];
          }
    if (!upload(frag.data(), frag.size() * sizeof(unsigned), reinterpret_cast<void**>(&c->dBFrag))) {
      b200sdr_chain_destroy(c);
      return st;
    }
  }
  if (c->hasAudioFir() && c->elem == kElemInt8Complex) {
    c->toepPlan = planToeplitz(c->T1, c->D1, c->mod, c->T2, c->D2, c->device);
    if (c->toepPlan.ok) {
      std::vector<uint32_t> frag;
      buildToeplitzFragments(cfg->rf_taps, c->T1, c->D1, c->mix, c->phaseStep, static_cast<double>(c->inScale), c->toepPlan, frag, c->toepScale);
      if (c->mix) c->rot1 = hostPhasor(c->phaseStep * static_cast<uint64_t>(c->D1));
      std::vector<float> taps2((static_cast<size_t>(c->T2) + 3u) & ~static_cast<size_t>(3), 0.0f);
      for (unsigned i = 0; i < c->T2; i++) taps2[i] = cfg->audio_taps[i];
      if (!upload(frag.data(), frag.size() * sizeof(uint32_t), reinterpret_cast<void**>(&c->dToepFrag)) ||
          !upload(taps2.data(), taps2.size() * sizeof(float), reinterpret_cast<void**>(&c->dToepTaps2))) {
        b200sdr_chain_destroy(c);
        return st;
      }
    }
  }
  if (c->toepPlan.ok) {
    c->variant = toeplitzVariantName(c->toepPlan, buf, sizeof(buf));
  } else if (c->fusedPlan.fused) {
    c->variant = chainVariantName(c->elem, c->mix, c->fusedPlan, buf, sizeof(buf));
  } else {
    c->variant = firVariantName(c->elem, false, c->mix, c->tableRoute, buf, sizeof(buf));
  }
  *chainOut = c;
  return B200SDR_OK;
}

B200SDR_EXPORT const char* b200sdr_chain_variant(const b200sdr_chain* c) { return c ? c->variant.c_str() : ""; }

B200SDR_EXPORT void b200sdr_chain_counts(const b200sdr_chain* c, size_t numInputs, size_t* numRf, size_t* numDemod, size_t* numAudio) {
  size_t rf = 0, demod = 0, audio = 0;
  if (c) {
    rf = firNumOutputs(numInputs, c->T1, c->D1);
    demod = c->mod == kModFm ? (rf == 0 ? 0 : rf - 1) : rf;  // QuadFmDemod.cpp:76-84
    audio = c->hasAudioFir() ? firNumOutputs(demod, c->T2, c->D2) : demod;
  }
  if (numRf) *numRf = rf;
  if (numDemod) *numDemod = demod;
  if (numAudio) *numAudio = audio;
}

B200SDR_EXPORT size_t b200sdr_chain_input_stride(const b200sdr_chain* c) { return c ? c->stride() : 0; }
B200SDR_EXPORT size_t b200sdr_chain_input_window(const b200sdr_chain* c) { return c ? c->window() : 0; }

B200SDR_EXPORT b200sdr_status b200sdr_chain_segment(
    const b200sdr_chain* c, size_t numAudio, size_t parts, size_t index, size_t* firstOutput, size_t* outputCount,
    size_t* firstInput, size_t* inputCount) {
  if (!c || parts == 0 || index >= parts) return fail(B200SDR_INVALID_ARGUMENT, "bad segment request");
  const size_t base = numAudio / parts, extra = numAudio % parts;
  const size_t first = index * base + (index < extra ? index : extra);
  const size_t count = base + (index < extra ? 1 : 0);
  if (firstOutput) *firstOutput = first;
  if (outputCount) *outputCount = count;
  if (firstInput) *firstInput = first * c->stride();
  if (inputCount) *inputCount = count == 0 ? 0 : (count - 1) * c->stride() + c->window();
  return B200SDR_OK;
}

B200SDR_EXPORT b200sdr_status b200sdr_chain_segment_weighted(
    const b200sdr_chain* c, size_t numAudio, size_t parts, const double* weights, size_t index, size_t* firstOutput, size_t* outputCount,
    size_t* firstInput, size_t* inputCount) {
  if (!weights) return b200sdr_chain_segment(c, numAudio, parts, index, firstOutput, outputCount, firstInput, inputCount);
  if (!c || parts == 0 || index >= parts) return fail(B200SDR_INVALID_ARGUMENT, "bad segment request");
  double total = 0.0, before = 0.0, through = 0.0;
  for (size_t i = 0; i < parts; i++) {
    if (!(weights[i] > 0.0)) return fail(B200SDR_INVALID_ARGUMENT, "segment weights must be positive");
    total += weights[i];
    if (i < index) before += weights[i];
    if (i <= index) through += weights[i];
  }
  // boundaries are rounded from the cumulative weights, so the parts tile [0, numAudio) exactly whatever the weights
  const size_t first = static_cast<size_t>(std::llround(static_cast<double>(numAudio) * (before / total)));
  const size_t end = index + 1 == parts ? numAudio : static_cast<size_t>(std::llround(static_cast<double>(numAudio) * (through / total)));
  const size_t count = end > first ? end - first : 0;
  if (firstOutput) *firstOutput = first;
  if (outputCount) *outputCount = count;
  if (firstInput) *firstInput = first * c->stride();
  if (inputCount) *inputCount = count == 0 ? 0 : (count - 1) * c->stride() + c->window();
  return B200SDR_OK;
}

// ---------------------------------------------------------------------------------------------------
B200SDR_EXPORT b200sdr_status b200sdr_chain_rf_stage(
    b200sdr_chain* c, const void* input, size_t numInputs, uint64_t firstSampleIndex, void* output,
    size_t numOutputs, cudaStream_t stream) {
  if (!c) return fail(B200SDR_INVALID_ARGUMENT, "chain is null");
  if (numOutputs == 0) return B200SDR_OK;
  if (!input || !output) return fail(B200SDR_INVALID_ARGUMENT, "input/output is null");
  const size_t needed = (numOutputs - 1 + c->fm()) * static_cast<size_t>(c->D1) + c->T1;
  if (numInputs < needed) return fail(B200SDR_OUT_OF_RANGE, "numOutputs needs more input samples than numInputs");
  DeviceGuard guard(c->device);
  if (guard.status != cudaSuccess) return cudaFail(guard.status, "cudaSetDevice");

  FirParams prm {};
  prm.in = input;
  prm.out = output;
  prm.taps = c->dTaps1;
  prm.nOut = numOutputs;
  prm.nIn = numInputs;
  prm.firstIndex = firstSampleIndex;
  prm.phaseStep = c->phaseStep;
  prm.T = c->T1;
  prm.D = c->D1;
  prm.mod = c->mod;
  prm.gain = c->fmGain;
  prm.inScale = c->inScale;
  if (c->tableRoute.rows) {  // used only if this call also routes to rows (same MP: depends on T, D only)
    prm.tapTable = c->dTapTable;
    prm.mixTable = c->dMixTable;
    prm.rotTable = c->dRotTable;
  }
  CUDA_OR_RETURN(launchFir(c->elem, false, c->mix, prm, stream));
  return B200SDR_OK;
}

B200SDR_EXPORT b200sdr_status b200sdr_chain_audio_stage(
    b200sdr_chain* c, const float* demod, float* audio, size_t numAudio, cudaStream_t stream) {
  if (!c) return fail(B200SDR_INVALID_ARGUMENT, "chain is null");
  if (!c->hasAudioFir()) return fail(B200SDR_INVALID_STATE, "chain has no audio FIR");
  if (numAudio == 0) return B200SDR_OK;
  if (!demod || !audio) return fail(B200SDR_INVALID_ARGUMENT, "demod/audio is null");
  DeviceGuard guard(c->device);
  if (guard.status != cudaSuccess) return cudaFail(guard.status, "cudaSetDevice");
  FirParams prm {};
  prm.in = demod;
  prm.out = audio;
  prm.taps = c->dTaps2;
  prm.nOut = numAudio;
  prm.T = c->T2;
  prm.D = c->D2;
  prm.nIn = (numAudio - 1) * static_cast<size_t>(c->D2) + c->T2;
  prm.mod = kModNone;
  prm.gain = 1.0f;
  prm.inScale = 1.0f;
  CUDA_OR_RETURN(launchFir(kElemReal, false, false, prm, stream));
  return B200SDR_OK;
}

B200SDR_EXPORT b200sdr_status b200sdr_chain_run(
    b200sdr_chain* c, const void* input, size_t numInputs, uint64_t firstSampleIndex, float* demodScratch, float* audio,
    size_t numAudio, cudaStream_t stream) {
  if (!c) return fail(B200SDR_INVALID_ARGUMENT, "chain is null");
  if (numAudio == 0) return B200SDR_OK;
  if (!input || !audio) return fail(B200SDR_INVALID_ARGUMENT, "input/audio is null");
  if (numInputs < (numAudio - 1) * c->stride() + c->window())
    return fail(B200SDR_OUT_OF_RANGE, "numAudio outputs need more input samples than numInputs");
  if (!c->hasAudioFir()) return b200sdr_chain_rf_stage(c, input, numInputs, firstSampleIndex, audio, numAudio, stream);

  if (c->toepPlan.ok && (reinterpret_cast<uintptr_t>(input) & 15u) == 0 && numInputs >= 128) {
    // one persistent kernel, RF stage on the int8 tensor cores (demodScratch is not touched)
    DeviceGuard guard(c->device);
    if (guard.status != cudaSuccess) return cudaFail(guard.status, "cudaSetDevice");
    ToepParams prm {};
    prm.in = static_cast<const unsigned char*>(input);
    prm.out = audio;
    prm.taps2 = c->dToepTaps2;
    prm.bFrag = c->dToepFrag;
    prm.nInBytes = static_cast<unsigned long long>(numInputs) * 2ull;
    prm.nAudio = numAudio;
    prm.D1 = c->D1;
    prm.T2 = c->T2;
    prm.D2 = c->D2;
    prm.fm = c->mod == kModFm ? 1 : 0;
    prm.gain = c->fmGain;
    prm.rot1 = c->rot1;
    prm.s0 = c->toepScale[0];
    prm.s1 = c->toepScale[1];
    prm.s2 = c->toepScale[2];
    CUDA_OR_RETURN(launchToeplitz(c->toepPlan, prm, stream));
    return B200SDR_OK;
  }
  if (c->fusedPlan.fused && (reinterpret_cast<uintptr_t>(input) & 15u) == 0) {
    // one persistent kernel: the demodulated stream stays in shared memory (demodScratch is not touched)
    DeviceGuard guard(c->device);
    if (guard.status != cudaSuccess) return cudaFail(guard.status, "cudaSetDevice");
    ChainParams prm {};
    prm.in = input;
    prm.out = audio;
    prm.tapTable = c->dTapTable;
    prm.mixTable = c->dMixTable;
    prm.rotTable = c->dRotTable;
    prm.taps2 = c->dTaps2;
    prm.bFrag = c->dBFrag;
    prm.kSteps = c->fusedPlan.kSteps;
    prm.digitScale[0] = c->digitScale[0];
    prm.digitScale[1] = c->digitScale[1];
    prm.digitScale[2] = c->digitScale[2];
    prm.nIn = numInputs;
    prm.nAudio = numAudio;
    prm.T1 = c->T1;
    prm.D1 = c->D1;
    prm.T2 = c->T2;
    prm.D2 = c->D2;
    prm.mod = c->mod;
    prm.gain = c->fmGain;
    CUDA_OR_RETURN(launchChain(c->elem, c->mix, c->fusedPlan, prm, stream));
    return B200SDR_OK;
  }
  // two kernels: K1 writes the demodulated stream to demodScratch, K2 (audio FIR) reads it back
  if (!demodScratch) return fail(B200SDR_INVALID_ARGUMENT, "demodScratch is null and this shape/alignment cannot take the fused kernel");
  const size_t nDemod = (numAudio - 1) * static_cast<size_t>(c->D2) + c->T2;
  b200sdr_status st = b200sdr_chain_rf_stage(c, input, numInputs, firstSampleIndex, demodScratch, nDemod, stream);
  if (st != B200SDR_OK) return st;
  return b200sdr_chain_audio_stage(c, demodScratch, audio, numAudio, stream);
}

B200SDR_EXPORT b200sdr_status b200sdr_chain_process_device(
    b200sdr_chain* c, const void* input, size_t numInputs, uint64_t firstSampleIndex, float* demodScratch,
    float* audio, size_t audioCapacity, size_t* numAudioOut, cudaStream_t stream) {
  if (numAudioOut) *numAudioOut = 0;
  if (!c) return fail(B200SDR_INVALID_ARGUMENT, "chain is null");
  size_t nRf, nDemod, nAudio;
  b200sdr_chain_counts(c, numInputs, &nRf, &nDemod, &nAudio);  // the reference's count rules (Fir.cpp:141-187)
  if (nAudio > audioCapacity) nAudio = audioCapacity;          // produce what fits; nothing is skipped (caller keeps input)
  if (nAudio == 0) return B200SDR_OK;
  const b200sdr_status st = b200sdr_chain_run(c, input, numInputs, firstSampleIndex, demodScratch, audio, nAudio, stream);
  if (st == B200SDR_OK && numAudioOut) *numAudioOut = nAudio;
  return st;
}

// ---------------------------------------------------------------------------------------------------
// Host-buffer path
// ---------------------------------------------------------------------------------------------------
B200SDR_EXPORT b200sdr_status b200sdr_chain_set_host_segment(b200sdr_chain* c, size_t inputSamples) {
  if (!c || inputSamples == 0) return fail(B200SDR_INVALID_ARGUMENT, "bad segment size");
  c->hostSegment = inputSamples;
  return B200SDR_OK;
}

static b200sdr_status ensureHostPath(b200sdr_chain* c, size_t inBytes, size_t demodCount, size_t audioCount) {
  if (!c->h2d) {
    CUDA_OR_RETURN(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
    CUDA_OR_RETURN(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
    CUDA_OR_RETURN(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
    for (int s = 0; s < kSlots; s++) {
      CUDA_OR_RETURN(cudaEventCreateWithFlags(&c->copied[s], cudaEventDisableTiming));
      CUDA_OR_RETURN(cudaEventCreateWithFlags(&c->computed[s], cudaEventDisableTiming));
      CUDA_OR_RETURN(cudaEventCreateWithFlags(&c->drained[s], cudaEventDisableTiming));
    }
  }
  if (inBytes > c->slotInBytes || demodCount > c->slotDemodCount || audioCount > c->slotAudioCount) {
    CUDA_OR_RETURN(cudaDeviceSynchronize());
    for (int s = 0; s < kSlots; s++) {
      if (c->slotIn[s]) cudaFree(c->slotIn[s]);
      if (c->slotDemod[s]) cudaFree(c->slotDemod[s]);
      if (c->slotAudio[s]) cudaFree(c->slotAudio[s]);
      c->slotIn[s] = nullptr;
      c->slotDemod[s] = c->slotAudio[s] = nullptr;
      CUDA_OR_RETURN(cudaMalloc(&c->slotIn[s], inBytes));
      CUDA_OR_RETURN(cudaMalloc(reinterpret_cast<void**>(&c->slotDemod[s]), sizeof(float) * (demodCount ? demodCount : 1)));
      CUDA_OR_RETURN(cudaMalloc(reinterpret_cast<void**>(&c->slotAudio[s]), sizeof(float) * (audioCount ? audioCount : 1)));
    }
    c->slotInBytes = inBytes;
    c->slotDemodCount = demodCount;
    c->slotAudioCount = audioCount;
  }
  return B200SDR_OK;
}

B200SDR_EXPORT b200sdr_status b200sdr_chain_process_host(
    b200sdr_chain* c, const void* hostInput, size_t numInputs, uint64_t firstSampleIndex, float* hostAudio,
    size_t audioCapacity, size_t* numAudioOut) {
  if (numAudioOut) *numAudioOut = 0;
  if (!c) return fail(B200SDR_INVALID_ARGUMENT, "chain is null");
  if (!hostInput || !hostAudio) return fail(B200SDR_INVALID_ARGUMENT, "hostInput/hostAudio is null");
  size_t nRf, nDemod, nAudio;
  b200sdr_chain_counts(c, numInputs, &nRf, &nDemod, &nAudio);
  if (nAudio > audioCapacity) nAudio = audioCapacity;
  if (nAudio == 0) return B200SDR_OK;
  DeviceGuard guard(c->device);
  if (guard.status != cudaSuccess) return cudaFail(guard.status, "cudaSetDevice");

  const size_t stride = c->stride(), window = c->window();
  const size_t outFloats = c->outElemFloats();
  size_t outPerSeg = c->hostSegment / stride;
  if (outPerSeg == 0) outPerSeg = 1;
  if (outPerSeg > nAudio) outPerSeg = nAudio;
  const size_t segInputs = (outPerSeg - 1) * stride + window;
  const size_t segDemod = c->hasAudioFir() ? (outPerSeg - 1) * c->D2 + c->T2 : 0;
  b200sdr_status st = ensureHostPath(c, segInputs * c->elemBytes(), segDemod, outPerSeg * outFloats);
  if (st != B200SDR_OK) return st;

  const unsigned char* hostBytes = static_cast<const unsigned char*>(hostInput);
  const size_t segments = (nAudio + outPerSeg - 1) / outPerSeg;
  for (size_t s = 0; s < segments; s++) {
    const int slot = static_cast<int>(s % kSlots);
    const size_t a0 = s * outPerSeg;
    const size_t count = nAudio - a0 < outPerSeg ? nAudio - a0 : outPerSeg;
    const size_t in0 = a0 * stride;
    const size_t inCount = (count - 1) * stride + window;
    // the slot's input may be overwritten once the kernels of its previous use have run, its audio
    // buffer once the previous D2H has drained
    if (s >= kSlots) {
      CUDA_OR_RETURN(cudaStreamWaitEvent(c->h2d, c->computed[slot], 0));
      CUDA_OR_RETURN(cudaStreamWaitEvent(c->compute, c->drained[slot], 0));
    }
    CUDA_OR_RETURN(cudaMemcpyAsync(c->slotIn[slot], hostBytes + in0 * c->elemBytes(), inCount * c->elemBytes(), cudaMemcpyHostToDevice, c->h2d));
    CUDA_OR_RETURN(cudaEventRecord(c->copied[slot], c->h2d));
    CUDA_OR_RETURN(cudaStreamWaitEvent(c->compute, c->copied[slot], 0));
    st = b200sdr_chain_run(c, c->slotIn[slot], inCount, firstSampleIndex + in0, c->slotDemod[slot], c->slotAudio[slot], count, c->compute);
    if (st != B200SDR_OK) return st;
    CUDA_OR_RETURN(cudaEventRecord(c->computed[slot], c->compute));
    CUDA_OR_RETURN(cudaStreamWaitEvent(c->d2h, c->computed[slot], 0));
    CUDA_OR_RETURN(cudaMemcpyAsync(hostAudio + a0 * outFloats, c->slotAudio[slot], count * outFloats * sizeof(float), cudaMemcpyDeviceToHost, c->d2h));
    CUDA_OR_RETURN(cudaEventRecord(c->drained[slot], c->d2h));
  }
  CUDA_OR_RETURN(cudaStreamSynchronize(c->d2h));
  CUDA_OR_RETURN(cudaStreamSynchronize(c->compute));
  if (numAudioOut) *numAudioOut = nAudio;
  return B200SDR_OK;
}
