// The wideband channelizer behind include/b200sdr/b200sdr.h: table construction (fp64 -> 24-bit fixed point ->
// IMMA fragment order), the RF kernel launch over (channel groups x row tiles) and the batched audio FIR.
#include <b200sdr/b200sdr.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "channel_kernels.cuh"
#include "channel_tc_kernels.cuh"
#include "digits.h"
#include "fir_dispatch.h"
#include "fir_kernels.cuh"
#include "pfb256_kernels.cuh"
#include "pfb_kernels.cuh"
#include "toeplitz_dispatch.h"

using namespace b200sdr;

namespace b200sdr {
b200sdr_status chainFail(b200sdr_status status, const std::string& what);  // chain.cu: sets b200sdr_last_error()
}

namespace {

size_t firCount(size_t nIn, size_t T, size_t D) { return (T == 0 || nIn + 1 < T) ? 0 : (nIn + 1 - T) / (D == 0 ? 1 : D); }

// Do all frequencies equal f[0] + b * fs / 2^log2N for integers b?  Decided on the 64-bit phase steps the kernels mix with
// (turns per sample, mod 1), with a slack of 2^-52 turns for their rounding; bins[ch] receives b mod 2^log2N.
bool onRaster(const double* frequencies, unsigned count, double sampleRate, unsigned log2N, int* bins) {
  const unsigned shift = 64u - log2N;
  const uint64_t step0 = phaseStepOf(frequencies[0], sampleRate);
  for (unsigned ch = 0; ch < count; ch++) {
    const uint64_t d = phaseStepOf(frequencies[ch], sampleRate) - step0;  // wraps: turns are mod 1
    const uint64_t b = (d + (1ull << (shift - 1))) >> shift;              // nearest bin
    const int64_t resid = static_cast<int64_t>(d - (b << shift));
    if (resid < -4096 || resid > 4096) return false;
    if (bins) bins[ch] = static_cast<int>(b & ((1ull << log2N) - 1ull));
  }
  return true;
}

b200sdr_status cudaFailC(cudaError_t e, const char* where) {
  return chainFail(e == cudaErrorMemoryAllocation ? B200SDR_OUT_OF_MEMORY : B200SDR_RUNTIME_ERROR, std::string(where) + ": " + cudaGetErrorString(e));
}

}  // namespace

struct b200sdr_channelizer {
  int device = 0;
  unsigned C = 0, T1 = 0, D1 = 1, M = 0, NTC = 1, KS = 0, T2 = 0, D2 = 1, groups = 0, warps = 4;
  bool anyFm = false, anyAm = false;
  float* dTail = nullptr;  // [C]: the AM channels' extra last output of a mixed AM/FM set (b200sdr_channelizer_process)
  unsigned* dBFrag = nullptr;
  float2* dRot = nullptr;
  float* dScale = nullptr;
  float* dGain = nullptr;
  int* dMod = nullptr;
  float* dTaps2 = nullptr;
  // tcgen05 route of the per-channel GEMM (channel_tc_kernels.cuh): B as [group][240 columns][K] int8, K-major
  bool tc = false;
  unsigned tcGroups = 0, tcSlabs = 0;
  signed char* dBtc = nullptr;
  CUtensorMap tcMapB {};
  // polyphase-filter-bank route (pfb_kernels.cuh): every channel on a raster fs/N with one common offset
  bool pfb = false;
  unsigned pfbN = 0, pfbQn = 0, pfbSmem = 0;
  double* dPfbTapsRe = nullptr;
  double* dPfbTapsIm = nullptr;
  double2* dPfbAcc0 = nullptr;
  double2* dPfbTwiddle = nullptr;
  int* dPfbBin = nullptr;
  int* dPfbOrder = nullptr;
  float2* dPfbRot1 = nullptr;
  // N = 256 specialisation (pfb256_kernels.cuh): taps in registers, register-resident FFT, two CTAs per SM
  unsigned pfb256QN = 0;  // 0: not taken
  float4* dPfb256Info = nullptr;
  std::vector<int> hostMods;
  std::string variant;
};

B200SDR_EXPORT void b200sdr_channelizer_destroy(b200sdr_channelizer* c) {
  if (!c) return;
  DeviceGuard guard(c->device);
  cudaFree(c->dTail);
  cudaFree(c->dBtc);
  cudaFree(c->dBFrag);
  cudaFree(c->dRot);
  cudaFree(c->dScale);
  cudaFree(c->dGain);
  cudaFree(c->dMod);
  cudaFree(c->dTaps2);
  cudaFree(c->dPfbTapsRe);
  cudaFree(c->dPfbTapsIm);
  cudaFree(c->dPfbAcc0);
  cudaFree(c->dPfbTwiddle);
  cudaFree(c->dPfbBin);
  cudaFree(c->dPfbOrder);
  cudaFree(c->dPfbRot1);
  cudaFree(c->dPfb256Info);
  delete c;
}

B200SDR_EXPORT const char* b200sdr_channelizer_variant(const b200sdr_channelizer* c) { return c ? c->variant.c_str() : ""; }

B200SDR_EXPORT uint32_t b200sdr_channelizer_raster(const double* frequencies, uint32_t numChannels, double sampleRate, int32_t* bins) {
  if (!frequencies || numChannels == 0 || !(sampleRate > 0.0)) return 0;
  for (unsigned log2N = 2; log2N <= 8; log2N++)
    if (onRaster(frequencies, numChannels, sampleRate, log2N, bins)) return 1u << log2N;
  return 0;
}

B200SDR_EXPORT b200sdr_status b200sdr_channelizer_create(const b200sdr_channelizer_config* cfg, b200sdr_channelizer** out) {
  if (!cfg || !out) return chainFail(B200SDR_INVALID_ARGUMENT, "config and out must be non-null");
  *out = nullptr;
  if (cfg->struct_size != sizeof(b200sdr_channelizer_config)) return chainFail(B200SDR_INVALID_ARGUMENT, "struct_size mismatch");
  if (cfg->num_channels == 0 || !cfg->frequencies || !cfg->modulations) return chainFail(B200SDR_INVALID_ARGUMENT, "channels are required");
  if (!cfg->rf_taps || cfg->rf_tap_count == 0 || !cfg->audio_taps || cfg->audio_tap_count == 0)
    return chainFail(B200SDR_INVALID_ARGUMENT, "RF and audio taps are required");
  if (!(cfg->sample_rate > 0.0)) return chainFail(B200SDR_INVALID_ARGUMENT, "sample_rate must be positive");
  const size_t D = cfg->rf_decimation == 0 ? 1 : cfg->rf_decimation;
  const size_t M = (cfg->rf_tap_count + D - 1) / D;
  if (D % 8 != 0 || D > (1u << 20)) return chainFail(B200SDR_INVALID_ARGUMENT, "rf_decimation must be a multiple of 8");
  if (M > 8) return chainFail(B200SDR_INVALID_ARGUMENT, "at most 8 taps per decimation phase (rf_tap_count <= 8 * rf_decimation)");
  for (uint32_t i = 0; i < cfg->num_channels; i++) {
    if (cfg->modulations[i] > B200SDR_MOD_FM) return chainFail(B200SDR_INVALID_ARGUMENT, "channel modulation must be AM or FM");
    if (cfg->modulations[i] == B200SDR_MOD_FM && !cfg->fm_gains) return chainFail(B200SDR_INVALID_ARGUMENT, "fm_gains is required for FM channels");
  }

  b200sdr_channelizer* c = new (std::nothrow) b200sdr_channelizer();
  if (!c) return chainFail(B200SDR_OUT_OF_MEMORY, "host allocation failed");
  c->device = cfg->cuda_device;
  c->C = cfg->num_channels;
  c->T1 = static_cast<unsigned>(cfg->rf_tap_count);
  c->D1 = static_cast<unsigned>(D);
  c->M = static_cast<unsigned>(M);
  c->NTC = M <= 4 ? 1u : 2u;
  c->KS = (2u * c->D1 + 31u) / 32u;
  c->T2 = static_cast<unsigned>(cfg->audio_tap_count);
  c->D2 = cfg->audio_decimation == 0 ? 1u : static_cast<unsigned>(cfg->audio_decimation);
  c->groups = (c->C + kChanNC - 1) / kChanNC;
  const char* w = std::getenv("B200SDR_CHANNEL_WARPS");
  c->warps = (w && std::atoi(w) == 8) ? 8u : 4u;

  // ---- tables -----------------------------------------------------------------------------------------------
  const unsigned K = c->KS * 32u, NCOL = c->NTC * 8u, NT = kChanNC * c->NTC;
  const size_t chunkWords = static_cast<size_t>(NT) * 3u * 64u;
  std::vector<unsigned> frag(static_cast<size_t>(c->groups) * c->KS * chunkWords, 0u);
  std::vector<float2> rot(static_cast<size_t>(c->C) * 8u, make_float2(1.0f, 0.0f));
  std::vector<float> scales(static_cast<size_t>(c->C) * 3u, 0.0f), gains(c->C, 1.0f);
  std::vector<int> mods(c->C, 0);
  std::vector<double> B(static_cast<size_t>(K) * NCOL);
  const double twoPi = 6.283185307179586476925286766559;
  // tcgen05 route: needs whole 128-byte K-slabs (K = 2 D1 a multiple of 128) and the tensor-map encoder of the driver
  const unsigned Ktc = 2u * c->D1;
  {
    const char* e = std::getenv("B200SDR_CHANNEL_TC");
    c->tc = !(e && std::atoi(e) == 0) && Ktc % 128u == 0;
  }
  c->tcGroups = (c->C + kTcChannels - 1) / kTcChannels;
  c->tcSlabs = Ktc / 128u;
  std::vector<signed char> Btc(c->tc ? static_cast<size_t>(c->tcGroups) * kTcN * Ktc : 0, 0);
  for (unsigned ch = 0; ch < c->C; ch++) {
    const uint64_t step = phaseStepOf(cfg->frequencies[ch], cfg->sample_rate);
    auto phasor = [&](uint64_t turns, double& re, double& im) {
      const double phi = twoPi * (static_cast<double>(static_cast<int64_t>(turns)) * (1.0 / 18446744073709551616.0));
      re = std::cos(phi);
      im = std::sin(phi);
    };
    std::fill(B.begin(), B.end(), 0.0);
    double bMax = 0.0;
    for (unsigned p = 0; p < c->D1; p++) {
      double wr, wi;
      phasor(step * p, wr, wi);
      wr *= 1.0 / 128.0;
      wi *= 1.0 / 128.0;
      for (unsigned m = 0; m < c->M; m++) {
        const size_t j = static_cast<size_t>(m) * c->D1 + p;
        const double h = j < c->T1 ? static_cast<double>(cfg->rf_taps[j]) : 0.0;
        const double gr = h * wr, gi = h * wi;
        B[static_cast<size_t>(2 * p) * NCOL + 2 * m] = gr;
        B[static_cast<size_t>(2 * p + 1) * NCOL + 2 * m] = -gi;
        B[static_cast<size_t>(2 * p) * NCOL + 2 * m + 1] = gi;
        B[static_cast<size_t>(2 * p + 1) * NCOL + 2 * m + 1] = gr;
      }
    }
    for (double v : B) bMax = std::fmax(bMax, std::fabs(v));
    const FixedPoint24 fx = fixedPoint24For(bMax);
    for (unsigned d = 0; d < 3; d++) scales[ch * 3u + d] = fx.digitScale[d];
    for (unsigned m = 0; m < 8; m++) {
      double re, im;
      phasor(step * (static_cast<uint64_t>(m) * c->D1), re, im);
      rot[ch * 8u + m] = make_float2(static_cast<float>(re), static_cast<float>(im));
    }
    mods[ch] = cfg->modulations[ch] == B200SDR_MOD_FM ? 1 : 0;
    c->anyFm = c->anyFm || mods[ch] == 1;
    c->anyAm = c->anyAm || mods[ch] == 0;
    gains[ch] = mods[ch] == 1 ? cfg->fm_gains[ch] : 1.0f;

    const unsigned group = ch / kChanNC, local = ch % kChanNC;
    for (unsigned ks = 0; ks < c->KS; ks++)
      for (unsigned nt = 0; nt < c->NTC; nt++)
        for (unsigned half = 0; half < 2; half++)
          for (unsigned lane = 0; lane < 32; lane++) {
            const unsigned g = lane >> 2, t = lane & 3u;
            unsigned word[3] = {0, 0, 0};
            for (unsigned e = 0; e < 4; e++) {
              const unsigned k = ks * 32u + half * 16u + t * 4u + e, n = nt * 8u + g;
              int dg[3];
              balancedDigits(B[static_cast<size_t>(k) * NCOL + n], fx.scale, dg);
              for (int d = 0; d < 3; d++) word[d] |= (static_cast<unsigned>(dg[d]) & 0xffu) << (8u * e);
              if (c->tc && k < Ktc) {  // the same digits, K-major: row = channel-in-group * 48 + digit * 16 + column
                const size_t rowBase = static_cast<size_t>(ch / kTcChannels) * kTcN + (ch % kTcChannels) * 48u + n;
                for (int d = 0; d < 3; d++) Btc[(rowBase + 16u * d) * Ktc + k] = static_cast<signed char>(dg[d]);
              }
            }
            const unsigned n = local * c->NTC + nt;
            for (unsigned d = 0; d < 3; d++)
              frag[(static_cast<size_t>(group) * c->KS + ks) * chunkWords + ((n * 3u + d) * 2u + half) * 32u + lane] = word[d];
          }
  }

  // ---- polyphase filter bank route: all channel frequencies = f0 + bin * fs / N for a power of two N <= 256 ----
  std::vector<double2> pfbTaps, pfbAcc0, pfbTwiddle;
  std::vector<int> pfbBin(c->C, 0);
  std::vector<float2> pfbRot1(c->C, make_float2(1.0f, 0.0f));
  {
    const char* e = std::getenv("B200SDR_PFB");
    const bool allowed = !(e && std::atoi(e) == 0) && c->C <= kPfbMaxN && c->D1 % 2u == 0;
    const uint64_t step0 = phaseStepOf(cfg->frequencies[0], cfg->sample_rate);  // the common offset goes into the taps
    for (unsigned log2N = 2; allowed && !c->pfb && log2N <= 8; log2N++) {
      if (!onRaster(cfg->frequencies, c->C, cfg->sample_rate, log2N, pfbBin.data())) continue;
      const unsigned N = 1u << log2N, Qn = (c->T1 + N - 1u) / N;
      const PfbSmem lay = pfbSmemLayout(N, Qn, c->D1, c->C);
      if (lay.total > 226u * 1024u) continue;  // a finer raster has fewer taps per phase: keep looking
      c->pfb = true;
      c->pfbN = N;
      c->pfbQn = Qn;
      c->pfbSmem = lay.total;
      pfbTaps.assign(static_cast<size_t>(Qn) * N, make_double2(0.0, 0.0));
      pfbAcc0.assign(N, make_double2(0.0, 0.0));
      pfbTwiddle.resize(N);
      for (unsigned j = 0; j < c->T1; j++) {
        const double phi = twoPi * (static_cast<double>(static_cast<int64_t>(step0 * j)) * (1.0 / 18446744073709551616.0));
        const double h = static_cast<double>(cfg->rf_taps[j]) * (1.0 / 128.0);
        pfbTaps[j] = make_double2(h * std::cos(phi), h * std::sin(phi));
      }
      const double magic = 1048576.0 + 128.0;  // the kernel feeds X = 2^20 + (x + 128) to the multiply-adds
      for (unsigned r = 0; r < N; r++) {
        double sr = 0.0, si = 0.0;
        for (unsigned q = 0; q < Qn; q++) {
          sr += pfbTaps[static_cast<size_t>(q) * N + r].x;
          si += pfbTaps[static_cast<size_t>(q) * N + r].y;
        }
        pfbAcc0[r] = make_double2(-magic * (sr - si), -magic * (sr + si));  // -(2^20 + 128) (1 + i) (sr + i si)
      }
      for (unsigned t = 0; t < N; t++) pfbTwiddle[t] = make_double2(std::cos(twoPi * t / N), std::sin(twoPi * t / N));
      for (unsigned ch = 0; ch < c->C; ch++) {
        const uint64_t step = phaseStepOf(cfg->frequencies[ch], cfg->sample_rate);
        const double phi = twoPi * (static_cast<double>(static_cast<int64_t>(step * static_cast<uint64_t>(c->D1))) * (1.0 / 18446744073709551616.0));
        pfbRot1[ch] = make_float2(static_cast<float>(std::cos(phi)), static_cast<float>(std::sin(phi)));
      }
    }
  }

  if (c->pfb) c->tc = false;  // the filter bank takes the launch: the GEMM tables of the tcgen05 route are not needed

  DeviceGuard guard(c->device);
  b200sdr_status st = B200SDR_OK;
  auto upload = [&](const void* host, size_t bytes, void** dev) -> bool {
    cudaError_t e = cudaMalloc(dev, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      st = cudaFailC(e, "uploading channelizer tables");
      return false;
    }
    return true;
  };
  bool ok = guard.status == cudaSuccess;
  if (!ok) st = cudaFailC(guard.status, "cudaSetDevice");
  ok = ok && upload(frag.data(), frag.size() * sizeof(unsigned), reinterpret_cast<void**>(&c->dBFrag)) &&
       upload(rot.data(), rot.size() * sizeof(float2), reinterpret_cast<void**>(&c->dRot)) &&
       upload(scales.data(), scales.size() * sizeof(float), reinterpret_cast<void**>(&c->dScale)) &&
       upload(gains.data(), gains.size() * sizeof(float), reinterpret_cast<void**>(&c->dGain)) &&
       upload(mods.data(), mods.size() * sizeof(int), reinterpret_cast<void**>(&c->dMod)) &&
       upload(cfg->audio_taps, sizeof(float) * c->T2, reinterpret_cast<void**>(&c->dTaps2));
  if (ok && c->tc) {
    const EncodeTiled encode = encodeTiled();
    if (!encode) {
      c->tc = false;
    } else {
      ok = upload(Btc.data(), Btc.size(), reinterpret_cast<void**>(&c->dBtc));
      const cuuint64_t dims[2] = {Ktc, static_cast<cuuint64_t>(c->tcGroups) * kTcN};
      const cuuint64_t strides[1] = {Ktc};
      const cuuint32_t box[2] = {128u, kTcN}, es[2] = {1u, 1u};
      if (ok && encode(&c->tcMapB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, c->dBtc, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        c->tc = false;
    }
  }
  if (ok && c->anyFm && c->anyAm) {
    const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->dTail), sizeof(float) * c->C);
    if (e != cudaSuccess) {
      st = cudaFailC(e, "cudaMalloc");
      ok = false;
    }
  }
  if (ok && c->pfb) {
    std::vector<int> pfbOrder;  // AM channels first, then FM: a warp demodulates 32 channels of one kind per instruction
    for (int kind = 0; kind < 2; kind++)
      for (unsigned ch = 0; ch < c->C; ch++)
        if (mods[ch] == kind) pfbOrder.push_back(static_cast<int>(ch));
    std::vector<double> re(pfbTaps.size()), im(pfbTaps.size());
    for (size_t i = 0; i < pfbTaps.size(); i++) {
      re[i] = pfbTaps[i].x;
      im[i] = pfbTaps[i].y;
    }
    ok = upload(re.data(), re.size() * sizeof(double), reinterpret_cast<void**>(&c->dPfbTapsRe)) &&
         upload(im.data(), im.size() * sizeof(double), reinterpret_cast<void**>(&c->dPfbTapsIm)) &&
         upload(pfbAcc0.data(), pfbAcc0.size() * sizeof(double2), reinterpret_cast<void**>(&c->dPfbAcc0)) &&
         upload(pfbTwiddle.data(), pfbTwiddle.size() * sizeof(double2), reinterpret_cast<void**>(&c->dPfbTwiddle)) &&
         upload(pfbBin.data(), pfbBin.size() * sizeof(int), reinterpret_cast<void**>(&c->dPfbBin)) &&
         upload(pfbOrder.data(), pfbOrder.size() * sizeof(int), reinterpret_cast<void**>(&c->dPfbOrder)) &&
         upload(pfbRot1.data(), pfbRot1.size() * sizeof(float2), reinterpret_cast<void**>(&c->dPfbRot1));
    // the N = 256 kernel: demodulation table in the same modulation-sorted order; needs one channel per bin at most
    const char* e256 = std::getenv("B200SDR_PFB256");
    unsigned qn = c->pfbQn <= 9u ? 9u : c->pfbQn <= 17u ? 17u : 0u;
    bool unique = true;
    {
      std::vector<char> seen(256, 0);
      for (unsigned ch = 0; ch < c->C && c->pfbN == 256u; ch++) {
        unique = unique && !seen[pfbBin[ch] & 255];
        seen[pfbBin[ch] & 255] = 1;
      }
    }
    if (ok && c->pfbN == 256u && qn != 0u && unique && pfb256Fits(c->D1, qn) && !(e256 && std::atoi(e256) == 0)) {
      std::vector<float4> info(256, make_float4(0.0f, 1.0f, 0.0f, 0.0f));
      for (size_t i = 0; i < pfbOrder.size(); i++) {
        const unsigned ch = static_cast<unsigned>(pfbOrder[i]);
        const unsigned bits = ch | (static_cast<unsigned>(pfbBin[ch]) & 255u) << 16 | (mods[ch] == 1 ? 1u << 24 : 0u) | 1u << 25;
        float w;
        std::memcpy(&w, &bits, sizeof(w));
        info[i] = make_float4(gains[ch], pfbRot1[ch].x, pfbRot1[ch].y, w);
      }
      ok = upload(info.data(), info.size() * sizeof(float4), reinterpret_cast<void**>(&c->dPfb256Info));
      if (ok) c->pfb256QN = qn;
    }
  }
  if (!ok) {
    b200sdr_channelizer_destroy(c);
    return st;
  }
  char buf[200];
  snprintf(buf, sizeof(buf), "channel<imma,NTC=%u>(channels=%u,groups=%u x %d,warps=%u,rowsTile=%u,kSteps=%u,M=%u) + batched direct FIR", c->NTC,
           c->C, c->groups, kChanNC, c->warps, c->warps * 32u, c->KS, c->M);
  if (c->tc)
    snprintf(buf, sizeof(buf), "channel<tcgen05,kind::i8>(channels=%u,groups=%u x %u,tile=128 rows x 240 columns,kSlabs=%u,M=%u,stages=%u,TMEM=2x256 columns) "
             "+ batched audio FIR", c->C, c->tcGroups, kTcChannels, c->tcSlabs, c->M, kTcStages);
  if (c->pfb)
    snprintf(buf, sizeof(buf), "pfb<N=%u,fp64>(channels=%u,taps/phase=%u,tile=%u RF outputs,warps=%u,smem=%u) + batched audio FIR", c->pfbN, c->C,
             c->pfbQn, kPfbTileK, kPfbWarps, c->pfbSmem);
  if (c->pfb256QN)
    snprintf(buf, sizeof(buf), "pfb<N=256,fp64>(kernel=pfb256<QN=%u>: taps in registers, register FFT 8x8x4, TMA input ring, 2 CTAs/SM x %u threads, "
             "channels=%u,taps/phase=%u,smem=%u) + batched audio FIR", c->pfb256QN, kP2Threads, c->C, c->pfbQn, pfb256SmemLayout().total);
  c->variant = buf;
  c->hostMods = mods;
  *out = c;
  return B200SDR_OK;
}

B200SDR_EXPORT void b200sdr_channelizer_counts(const b200sdr_channelizer* c, size_t numInputs, size_t* numDemod, size_t* numAudio) {
  size_t demod = 0, audio = 0;
  if (c) {
    const size_t rf = firCount(numInputs, c->T1, c->D1);
    // every channel uses the FM-safe count (one RF output held back) only if it is FM; per-channel counts differ by one,
    // so the common count is the smaller one when any channel is FM
    demod = c->anyFm ? (rf == 0 ? 0 : rf - 1) : rf;
    audio = firCount(demod, c->T2, c->D2);
  }
  if (numDemod) *numDemod = demod;
  if (numAudio) *numAudio = audio;
}

B200SDR_EXPORT b200sdr_status b200sdr_channelizer_channel_counts(
    const b200sdr_channelizer* c, uint32_t channel, size_t numInputs, size_t* numDemod, size_t* numAudio) {
  if (numDemod) *numDemod = 0;
  if (numAudio) *numAudio = 0;
  if (!c || channel >= c->C) return chainFail(B200SDR_INVALID_ARGUMENT, "channel out of range");
  const size_t rf = firCount(numInputs, c->T1, c->D1);
  const size_t demod = c->hostMods[channel] == 1 ? (rf == 0 ? 0 : rf - 1) : rf;  // QuadFmDemod.cpp:76-84 keeps one sample
  if (numDemod) *numDemod = demod;
  if (numAudio) *numAudio = firCount(demod, c->T2, c->D2);
  return B200SDR_OK;
}

B200SDR_EXPORT b200sdr_status b200sdr_channelizer_segment(
    const b200sdr_channelizer* c, size_t numAudio, size_t parts, size_t index, size_t* firstOutput, size_t* outputCount, size_t* firstInput,
    size_t* inputCount) {
  if (!c || parts == 0 || index >= parts) return chainFail(B200SDR_INVALID_ARGUMENT, "bad segment request");
  const size_t base = numAudio / parts, extra = numAudio % parts;
  const size_t first = index * base + (index < extra ? index : extra), count = base + (index < extra ? 1 : 0);
  const size_t stride = static_cast<size_t>(c->D1) * c->D2;
  const size_t window = (static_cast<size_t>(c->T2) - 1 + (c->anyFm ? 1 : 0)) * c->D1 + c->T1;  // RF taps + the demod samples of one audio window
  if (firstOutput) *firstOutput = first;
  if (outputCount) *outputCount = count;
  if (firstInput) *firstInput = first * stride;
  if (inputCount) *inputCount = count == 0 ? 0 : (count - 1) * stride + window;
  return B200SDR_OK;
}

namespace {

// the AM channels' rows of tail[] -> column `col` of audio (one thread per channel)
__global__ void scatterAmTail(const float* tail, const int* mod, float* audio, size_t audioStride, size_t col, unsigned channels) {
  const unsigned ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch < channels && mod[ch] == 0) audio[static_cast<size_t>(ch) * audioStride + col] = tail[ch];
}

b200sdr_status runImpl(b200sdr_channelizer* c, const void* input, size_t numInputs, float* demodScratch, size_t demodStride, float* audio,
                       size_t audioStride, size_t numAudio, bool forceAm, cudaStream_t stream);

}  // namespace

B200SDR_EXPORT b200sdr_status b200sdr_channelizer_run(
    b200sdr_channelizer* c, const void* input, size_t numInputs, float* demodScratch, size_t demodStride, float* audio, size_t audioStride,
    size_t numAudio, cudaStream_t stream) {
  return runImpl(c, input, numInputs, demodScratch, demodStride, audio, audioStride, numAudio, false, stream);
}

B200SDR_EXPORT b200sdr_status b200sdr_channelizer_process(
    b200sdr_channelizer* c, const void* input, size_t numInputs, float* demodScratch, size_t demodStride, float* audio, size_t audioStride,
    size_t* numAudioPerChannel, cudaStream_t stream) {
  if (!c) return chainFail(B200SDR_INVALID_ARGUMENT, "channelizer is null");
  size_t common = 0;
  b200sdr_channelizer_counts(c, numInputs, nullptr, &common);
  const size_t rf = firCount(numInputs, c->T1, c->D1);
  const size_t amAudio = firCount(rf, c->T2, c->D2);  // an AM channel's own count; the FM count is `common` when FM channels exist
  if (numAudioPerChannel)
    for (unsigned ch = 0; ch < c->C; ch++) numAudioPerChannel[ch] = c->hostMods[ch] == 1 ? common : amAudio;
  b200sdr_status st = runImpl(c, input, numInputs, demodScratch, demodStride, audio, audioStride, common, false, stream);
  if (st != B200SDR_OK || !(c->anyFm && c->anyAm) || amAudio == common) return st;
  // Mixed set and the AM channels own one more output than the FM ones (their demodulator holds no sample back): the AM
  // pass over the input window of that last output alone, every channel demodulated as AM, AM rows scattered into place.
  if (audioStride < amAudio || demodStride < c->T2) return chainFail(B200SDR_INVALID_ARGUMENT, "demodStride/audioStride too small for the AM channels");
  const size_t first = (amAudio - 1) * static_cast<size_t>(c->D1) * c->D2;  // a multiple of 8 samples: the slice stays 16-byte aligned
  st = runImpl(c, static_cast<const unsigned char*>(input) + 2 * first, numInputs - first, demodScratch, demodStride, c->dTail, 1, 1, true, stream);
  if (st != B200SDR_OK) return st;
  DeviceGuard guard(c->device);
  scatterAmTail<<<(c->C + 127u) / 128u, 128, 0, stream>>>(c->dTail, c->dMod, audio, audioStride, amAudio - 1, c->C);
  const cudaError_t e = launchStatus();
  if (e != cudaSuccess) return cudaFailC(e, "scatterAmTail launch");
  return B200SDR_OK;
}

namespace {

b200sdr_status runImpl(b200sdr_channelizer* c, const void* input, size_t numInputs, float* demodScratch, size_t demodStride, float* audio,
                       size_t audioStride, size_t numAudio, bool forceAm, cudaStream_t stream) {
  if (!c) return chainFail(B200SDR_INVALID_ARGUMENT, "channelizer is null");
  if (numAudio == 0) return B200SDR_OK;
  if (!input || !demodScratch || !audio) return chainFail(B200SDR_INVALID_ARGUMENT, "input/demodScratch/audio is null");
  if ((reinterpret_cast<uintptr_t>(input) & 15u) != 0) return chainFail(B200SDR_INVALID_ARGUMENT, "input must be 16-byte aligned");
  const size_t nDemod = (numAudio - 1) * static_cast<size_t>(c->D2) + c->T2;
  if (demodStride < nDemod || audioStride < numAudio) return chainFail(B200SDR_INVALID_ARGUMENT, "demodStride/audioStride too small");
  // demod output k reads rows k .. k+M (FM: one more RF output), i.e. input up to (k + 1) * D1 + T1 - 1 at most
  const size_t needed = (nDemod - 1 + 1) * static_cast<size_t>(c->D1) + c->T1;
  const size_t neededAm = (nDemod - 1) * static_cast<size_t>(c->D1) + c->T1;
  const bool anyFm = c->anyFm && !forceAm;
  if (numInputs < (anyFm ? needed : neededAm)) return chainFail(B200SDR_OUT_OF_RANGE, "numAudio outputs need more input samples than numInputs");
  DeviceGuard guard(c->device);
  if (guard.status != cudaSuccess) return cudaFailC(guard.status, "cudaSetDevice");

  cudaError_t e = cudaSuccess;
  if (c->pfb256QN) {
    Pfb256Params pp {};
    pp.in = static_cast<const unsigned char*>(input);
    pp.out = demodScratch;
    pp.tapsRe = c->dPfbTapsRe;
    pp.tapsIm = c->dPfbTapsIm;
    pp.acc0 = c->dPfbAcc0;
    pp.twiddle = c->dPfbTwiddle;
    pp.chanInfo = c->dPfb256Info;
    pp.nInBytes = static_cast<unsigned long long>(numInputs) * 2ull;
    pp.nOut = nDemod;
    pp.outStride = demodStride;
    pp.D1 = c->D1;
    pp.Qn = c->pfbQn;
    pp.C = c->C;
    pp.anyFm = anyFm ? 1 : 0;
    pp.forceAm = forceAm ? 1 : 0;
    auto kernel = c->pfb256QN == 9u ? pfb256Kernel<9> : pfb256Kernel<17>;
    const unsigned smem = pfb256SmemLayout().total;
    e = cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cudaFailC(e, "cudaFuncSetAttribute");
    int sms = kSmCount;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const unsigned long long units = (nDemod + 7) / 8;
    unsigned grid = 2u * static_cast<unsigned>(sms);
    const char* eg = std::getenv("B200SDR_PFB256_GRID");  // tests: few CTAs, so that each one walks many blocks
    if (eg && std::atoi(eg) > 0) grid = static_cast<unsigned>(std::atoi(eg));
    if (units < grid) grid = static_cast<unsigned>(units);
    kernel<<<grid, kP2Threads, smem, stream>>>(pp);
    e = launchStatus();
    if (e != cudaSuccess) return cudaFailC(e, "pfb256Kernel launch");
  } else if (c->pfb) {
    PfbParams pp {};
    pp.in = static_cast<const unsigned char*>(input);
    pp.out = demodScratch;
    pp.tapsRe = c->dPfbTapsRe;
    pp.tapsIm = c->dPfbTapsIm;
    pp.acc0 = c->dPfbAcc0;
    pp.twiddle = c->dPfbTwiddle;
    pp.bin = c->dPfbBin;
    pp.order = c->dPfbOrder;
    pp.mod = c->dMod;
    pp.gain = c->dGain;
    pp.rot1 = c->dPfbRot1;
    pp.nInBytes = static_cast<unsigned long long>(numInputs) * 2ull;
    pp.nOut = nDemod;
    pp.outStride = demodStride;
    pp.D1 = c->D1;
    pp.N = c->pfbN;
    pp.Qn = c->pfbQn;
    pp.C = c->C;
    pp.anyFm = anyFm ? 1 : 0;
    pp.forceAm = forceAm ? 1 : 0;
    e = cudaFuncSetAttribute(reinterpret_cast<const void*>(pfbKernel), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cudaFailC(e, "cudaFuncSetAttribute");
    int sms = kSmCount;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const unsigned tileOut = kPfbTileK - (anyFm ? 1u : 0u);
    const unsigned long long pfbTiles = (nDemod + tileOut - 1) / tileOut;
    const unsigned grid = pfbTiles < static_cast<unsigned long long>(sms) ? static_cast<unsigned>(pfbTiles) : static_cast<unsigned>(sms);
    pfbKernel<<<grid, kPfbWarps * 32u, c->pfbSmem, stream>>>(pp);
    e = launchStatus();
    if (e != cudaSuccess) return cudaFailC(e, "pfbKernel launch");
  } else {
  // ---- tcgen05 route: every output whose rows are complete in the input; the few that need the input's last, partial
  // row (zero-filled by the legacy kernel's loads) stay on the legacy kernel below ----
  size_t doneTc = 0;
  if (c->tc) {
    const size_t rowsAvail = numInputs / c->D1;
    const size_t fit = rowsAvail + (anyFm ? 0 : 1) > c->M ? rowsAvail + (anyFm ? 0 : 1) - c->M : 0;
    const size_t nTc = fit < nDemod ? fit : nDemod;
    const EncodeTiled encode = encodeTiled();
    if (nTc > 0 && encode) {
      const unsigned Ktc = 2u * c->D1;
      CUtensorMap mapA;
      const cuuint64_t dims[2] = {Ktc, rowsAvail};
      const cuuint64_t strides[1] = {Ktc};
      const cuuint32_t box[2] = {128u, kTcRows}, es[2] = {1u, 1u};
      if (encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(input), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return chainFail(B200SDR_RUNTIME_ERROR, "cuTensorMapEncodeTiled failed for the channelizer input");
      ChannelTcParams tp {};
      tp.out = demodScratch;
      tp.rot = c->dRot;
      tp.digitScale = c->dScale;
      tp.gain = c->dGain;
      tp.mod = c->dMod;
      tp.nOut = nTc;
      tp.outStride = demodStride;
      tp.M = c->M;
      tp.kSlabs = c->tcSlabs;
      tp.numChannels = c->C;
      tp.groups = c->tcGroups;
      tp.tiles = (nTc + (kTcRows - c->M) - 1) / (kTcRows - c->M);
      tp.forceAm = forceAm ? 1 : 0;
      const unsigned smem = channelTcSmemLayout(c->M).total;
      e = cudaFuncSetAttribute(reinterpret_cast<const void*>(channelTcKernel), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cudaFailC(e, "cudaFuncSetAttribute");
      int sms = kSmCount;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
      const unsigned long long items = tp.tiles * tp.groups;
      const unsigned grid = items < static_cast<unsigned long long>(sms) ? static_cast<unsigned>(items) : static_cast<unsigned>(sms);
      channelTcKernel<<<grid, kTcThreads, smem, stream>>>(tp, mapA, c->tcMapB);
      e = launchStatus();
      if (e != cudaSuccess) return cudaFailC(e, "channelTcKernel launch");
      doneTc = nTc;
    }
  }
  if (doneTc < nDemod) {
  ChannelParams prm {};
  prm.in = static_cast<const unsigned char*>(input) + 2 * doneTc * c->D1;
  prm.out = demodScratch + doneTc;
  prm.bFrag = c->dBFrag;
  prm.rot = c->dRot;
  prm.digitScale = c->dScale;
  prm.gain = c->dGain;
  prm.mod = c->dMod;
  prm.nIn = numInputs - doneTc * c->D1;
  prm.nOut = nDemod - doneTc;
  prm.outStride = demodStride;
  prm.D1 = c->D1;
  prm.M = c->M;
  prm.kSteps = c->KS;
  prm.numChannels = c->C;
  prm.forceAm = forceAm ? 1 : 0;
  const unsigned rowsTile = c->warps * 32u, OT = rowsTile - c->M;
  const unsigned NT = kChanNC * c->NTC;
  const size_t ring = static_cast<size_t>(kChanStages) * (static_cast<size_t>(rowsTile) * kChanARow + static_cast<size_t>(NT) * 3u * 64u * 4u);
  const size_t park = static_cast<size_t>(kChanNC) * c->M * rowsTile * sizeof(float2);
  const size_t smem = ring > park ? ring : park;
  const unsigned long long tiles = (nDemod - doneTc + OT - 1) / OT;
  if (tiles > 65535ull * 32768ull) return chainFail(B200SDR_OUT_OF_RANGE, "block too long");
  auto kernel = c->NTC == 1 ? channelKernel<1> : channelKernel<2>;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cudaFailC(e, "cudaFuncSetAttribute");
  }
  if (tiles > 65535ull) return chainFail(B200SDR_OUT_OF_RANGE, "more than 65535 row tiles in one call: split the block");
  kernel<<<dim3(c->groups, static_cast<unsigned>(tiles)), c->warps * 32u, smem, stream>>>(prm);
  e = launchStatus();
  if (e != cudaSuccess) return cudaFailC(e, "channelKernel launch");
  }
  }

  FirParams fir {};
  fir.in = demodScratch;
  fir.out = audio;
  fir.taps = c->dTaps2;
  fir.nOut = numAudio;
  fir.T = c->T2;
  fir.D = c->D2;
  fir.nIn = nDemod;
  fir.mod = kModNone;
  fir.gain = 1.0f;
  fir.inScale = 1.0f;
  fir.inBatchStride = demodStride;
  fir.outBatchStride = audioStride;
  // many taps per kept output (273 / 5 for C5): the register-tiled window kernel; else the thread-per-output kernel
  static const bool useWindow = !(std::getenv("B200SDR_AUDIO_WINDOW") && std::atoi(std::getenv("B200SDR_AUDIO_WINDOW")) == 0);
  fir.M = (fir.T + fir.D - 1) / fir.D;
  if (useWindow && windowEligible(kElemReal, false, false, fir))
    e = launchWindowBatched(kElemReal, fir, c->C, stream);
  else
    e = launchFirBatched(kElemReal, fir, c->C, stream);
  if (e != cudaSuccess) return cudaFailC(e, "batched audio FIR launch");
  return B200SDR_OK;
}

}  // namespace
