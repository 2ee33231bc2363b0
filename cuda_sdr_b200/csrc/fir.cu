// Host-side dispatch of the decimating-FIR kernels and the FIR entry points of the gsdr C-ABI.
#include <gsdr/gsdr.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <tuple>

#include "fir_dispatch.h"
#include "fir_kernels.cuh"

namespace b200sdr {

namespace {

using Kernel = void (*)(const FirParams);

// rows kernels are instantiated in fir_rows_*.cu (one translation unit per element type / mixer flag so
// that nvcc compiles them in parallel); [MP-1][RPT index: 1 row, high]
Kernel rowsKernelFor(int elem, bool mix, unsigned MP, int rptIdx) {
  const Kernel* table = elem == kElemInt8Complex ? (mix ? kRowsInt8Mix : kRowsInt8Plain) : (mix ? kRowsCf32Mix : kRowsCf32Plain);
  return table[(MP - 1) * 2 + rptIdx];
}

constexpr unsigned kMaxDynSmem = 200 * 1024;

int envInt(const char* name, int fallback) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : fallback;
}
// the routing knobs are read once per process (they exist for experiments, not for switching routes between two calls)
struct FirEnv {
  int forceDirect = envInt("B200SDR_FORCE_DIRECT", 0), rpt = envInt("B200SDR_RPT", 0), noWindow = envInt("B200SDR_NO_WINDOW", 0),
      noWideRows = envInt("B200SDR_NO_WIDE_ROWS", 0);
};
const FirEnv& firEnv() {
  static const FirEnv env;
  return env;
}

template <bool STAGED>
Kernel directKernelForT(int elem, bool tapc, bool mix) {
  switch (elem) {
    case kElemInt8Complex:
      return mix ? directKernel<kElemInt8Complex, false, true, false, STAGED> : directKernel<kElemInt8Complex, false, false, false, STAGED>;
    case kElemComplex:
      if (tapc) return directKernel<kElemComplex, true, false, false, STAGED>;
      return mix ? directKernel<kElemComplex, false, true, false, STAGED> : directKernel<kElemComplex, false, false, false, STAGED>;
    default:
      return tapc ? directKernel<kElemReal, true, false, false, STAGED> : directKernel<kElemReal, false, false, true, STAGED>;
  }
}

Kernel directKernelFor(int elem, bool tapc, bool mix, bool staged) {
  return staged ? directKernelForT<true>(elem, tapc, mix) : directKernelForT<false>(elem, tapc, mix);
}

}  // namespace

cudaError_t ensureDynamicSmem(const void* kernel, int bytes) {
  static std::mutex mutex;
  static std::set<std::tuple<int, const void*, int>> done;
  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mutex);
  if (done.count({device, kernel, bytes})) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.insert({device, kernel, bytes});
  return e;
}

FirRoute planFir(int elem, bool tapsComplex, const void* in, unsigned T, unsigned D, int mod) {
  FirRoute r {};
  r.rows = false;
  r.M = D ? (T + D - 1) / D : 0;
  const bool forceDirect = firEnv().forceDirect != 0;
  if (forceDirect || tapsComplex || elem == kElemReal || T == 0 || D == 0) return r;
  const unsigned es = elem == kElemInt8Complex ? 2u : 8u;
  const unsigned vec = 16u / es;
  if (D % vec != 0 || (reinterpret_cast<uintptr_t>(in) & 15u) != 0 || r.M > 8) return r;
  r.MP = r.M;
  r.TS = static_cast<unsigned>(tapStride(static_cast<int>(r.MP)));
  const int forcedRpt = firEnv().rpt;
  const unsigned fm = mod == kModFm ? 1u : 0u;
  for (int rptIdx = 1; rptIdx >= 0; rptIdx--) {
    const unsigned rpt = rptIdx ? rowsRptHigh(r.MP) : 1u;
    if (forcedRpt == 1 && rptIdx == 1) continue;
    const unsigned rowsPerTile = rpt * kRowsThreads;
    if (rowsPerTile <= r.M - 1 + fm) continue;
    const RowsSmem lay = rowsSmemLayout(D, r.TS, r.M, rowsPerTile, es, fm != 0);
    // keep at least ~3 tiles resident per SM for the high-RPT variant so TMA latency hides behind compute
    const unsigned limit = rptIdx ? 72u * 1024u : kMaxDynSmem;
    if (lay.total > limit) continue;
    r.rows = true;
    r.rptIdx = rptIdx;
    r.rpt = rpt;
    r.rowsPerTile = rowsPerTile;
    r.outPerTile = rowsPerTile - (r.M - 1) - fm;
    r.smemBytes = lay.total;
    return r;
  }
  return r;
}

cudaError_t launchFir(int elem, bool tapsComplex, bool mix, FirParams prm, cudaStream_t stream) {
  if (prm.nOut == 0) return cudaSuccess;
  if (prm.D == 0) prm.D = 1;
  const FirRoute route = planFir(elem, tapsComplex, prm.in, prm.T, prm.D, prm.mod);
  prm.M = route.M;
  if (route.rows) {
    prm.rowsPerTile = route.rowsPerTile;
    prm.outPerTile = route.outPerTile;
    const Kernel k = rowsKernelFor(elem, mix, route.MP, route.rptIdx);
    if (route.smemBytes > 48 * 1024) {
      const cudaError_t e = ensureDynamicSmem(reinterpret_cast<const void*>(k), static_cast<int>(kMaxDynSmem));
      if (e != cudaSuccess) return e;
    }
    const unsigned long long blocks = (prm.nOut + route.outPerTile - 1) / route.outPerTile;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    k<<<static_cast<unsigned>(blocks), kRowsThreads, route.smemBytes, stream>>>(prm);
    return launchStatus();
  }
  // 8 < M <= 32 on long rows (D >= 16; complex float data, real taps, no mixer): the rows kernel with 16 or 32 partial sums per
  // row.  Few outputs per input sample, so the cell is HBM-bound and wants the tile streamed ONCE by TMA; the window kernel would
  // walk the D phases in D / 4 strided passes.
  if (elem == kElemComplex && !tapsComplex && !mix && route.M > 8 && route.M <= 32 && prm.D >= 16 && prm.D % 2 == 0 &&
      (reinterpret_cast<uintptr_t>(prm.in) & 15u) == 0 && firEnv().noWideRows == 0) {
    const unsigned MP = route.M <= 16 ? 16u : 32u, fm = prm.mod == kModFm ? 1u : 0u;
    const unsigned rpt = (MP == 16 && 2u * kRowsThreads * rowsRowStride(prm.D, 8) <= 64u * 1024u) ? 2u : 1u;
    const unsigned rowsPerTile = rpt * kRowsThreads;
    const RowsSmem lay = rowsSmemLayout(prm.D, MP, route.M, rowsPerTile, 8, fm != 0);
    if (lay.total <= 100u * 1024u) {
      prm.rowsPerTile = rowsPerTile;
      prm.outPerTile = rowsPerTile - (route.M - 1) - fm;
      const Kernel k = kRowsCf32PlainWide[MP == 16 ? (rpt == 2 ? 1 : 0) : 2];
      if (lay.total > 48 * 1024) {
        const cudaError_t e = ensureDynamicSmem(reinterpret_cast<const void*>(k), static_cast<int>(kMaxDynSmem));
        if (e != cudaSuccess) return e;
      }
      const unsigned long long blocks = (prm.nOut + prm.outPerTile - 1) / prm.outPerTile;
      if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
      k<<<static_cast<unsigned>(blocks), kRowsThreads, lay.total, stream>>>(prm);
      return launchStatus();
    }
  }
  if (firEnv().noWindow == 0 && windowEligible(elem, tapsComplex, mix, prm)) return launchWindow(elem, prm, stream);
  const unsigned outPerBlock = prm.mod == kModFm ? kDirectThreads - 1 : kDirectThreads;
  const unsigned long long blocks = (prm.nOut + outPerBlock - 1) / outPerBlock;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  const unsigned long long elemBytes = elem == kElemInt8Complex ? 2 : elem == kElemComplex ? 8 : 4;
  const unsigned long long tileBytes = (static_cast<unsigned long long>(kDirectThreads - 1) * prm.D + prm.T) * elemBytes;
  const bool staged = tileBytes <= 96 * 1024;
  const unsigned smemBytes = kDirectFixedSmem + (staged ? static_cast<unsigned>(tileBytes) : 0u);
  const Kernel k = directKernelFor(elem, tapsComplex, mix, staged);
  if (smemBytes > 48 * 1024) {
    const cudaError_t e = ensureDynamicSmem(reinterpret_cast<const void*>(k), static_cast<int>(kMaxDynSmem));
    if (e != cudaSuccess) return e;
  }
  k<<<static_cast<unsigned>(blocks), kDirectThreads, smemBytes, stream>>>(prm);
  return launchStatus();
}

cudaError_t launchFirBatched(int elem, FirParams prm, unsigned batch, cudaStream_t stream) {
  if (prm.nOut == 0 || batch == 0) return cudaSuccess;
  if (prm.D == 0) prm.D = 1;
  if (batch > 65535u) return cudaErrorInvalidConfiguration;
  prm.M = (prm.T + prm.D - 1) / prm.D;
  const unsigned outPerBlock = prm.mod == kModFm ? kDirectThreads - 1 : kDirectThreads;
  const unsigned long long blocks = (prm.nOut + outPerBlock - 1) / outPerBlock;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  const unsigned long long elemBytes = elem == kElemInt8Complex ? 2 : elem == kElemComplex ? 8 : 4;
  const unsigned long long tileBytes = (static_cast<unsigned long long>(kDirectThreads - 1) * prm.D + prm.T) * elemBytes;
  const bool staged = tileBytes <= 96 * 1024;
  const unsigned smemBytes = kDirectFixedSmem + (staged ? static_cast<unsigned>(tileBytes) : 0u);
  const Kernel k = directKernelFor(elem, false, false, staged);
  if (smemBytes > 48 * 1024) {
    const cudaError_t e = ensureDynamicSmem(reinterpret_cast<const void*>(k), static_cast<int>(kMaxDynSmem));
    if (e != cudaSuccess) return e;
  }
  k<<<dim3(static_cast<unsigned>(blocks), batch), kDirectThreads, smemBytes, stream>>>(prm);
  return launchStatus();
}

const char* firVariantName(int elem, bool tapsComplex, bool mix, const FirRoute& route, char* buf, size_t bufLen) {
  const char* e = elem == kElemInt8Complex ? "int8c" : elem == kElemComplex ? "cf32" : "f32";
  if (route.rows) {
    snprintf(buf, bufLen, "rows<%s,mix=%d,MP=%u,RPT=%u>(rowsPerTile=%u,smem=%u)", e, mix ? 1 : 0, route.MP, route.rpt, route.rowsPerTile, route.smemBytes);
  } else {
    snprintf(buf, bufLen, "direct<%s,tapc=%d,mix=%d>", e, tapsComplex ? 1 : 0, mix ? 1 : 0);
  }
  return buf;
}

}  // namespace b200sdr

using namespace b200sdr;

static cudaError_t firEntry(
    int elem, bool tapc, size_t decimation, const void* taps, size_t tapCount, const void* input, void* output,
    size_t numOutputs, int32_t cudaDevice, cudaStream_t cudaStream) {
  DeviceGuard guard(cudaDevice);
  if (guard.status != cudaSuccess) return guard.status;
  if (numOutputs == 0) return cudaSuccess;
  if (tapCount == 0 || tapCount > 0xffffffffull || decimation > 0xffffffffull) return cudaErrorInvalidValue;
  FirParams prm {};
  prm.in = input;
  prm.out = output;
  prm.taps = static_cast<const float*>(taps);
  prm.nOut = numOutputs;
  prm.T = static_cast<unsigned>(tapCount);
  prm.D = decimation == 0 ? 1u : static_cast<unsigned>(decimation);
  prm.nIn = (numOutputs - 1) * static_cast<unsigned long long>(prm.D) + tapCount;
  prm.mod = kModNone;
  prm.gain = 1.0f;
  prm.inScale = 1.0f;
  return launchFir(elem, tapc, false, prm, cudaStream);
}

GSDR_EXPORT cudaError_t gsdrFirFF(
    size_t decimation, const float* taps, size_t tapCount, const float* input, float* output, size_t numOutputs,
    int32_t cudaDevice, cudaStream_t cudaStream) {
  return firEntry(kElemReal, false, decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream);
}

GSDR_EXPORT cudaError_t gsdrFirFC(
    size_t decimation, const float* taps, size_t tapCount, const cuComplex* input, cuComplex* output,
    size_t numOutputs, int32_t cudaDevice, cudaStream_t cudaStream) {
  return firEntry(kElemComplex, false, decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream);
}

GSDR_EXPORT cudaError_t gsdrFirCC(
    size_t decimation, const cuComplex* taps, size_t tapCount, const cuComplex* input, cuComplex* output,
    size_t numOutputs, int32_t cudaDevice, cudaStream_t cudaStream) {
  return firEntry(kElemComplex, true, decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream);
}

GSDR_EXPORT cudaError_t gsdrFirCF(
    size_t decimation, const cuComplex* taps, size_t tapCount, const float* input, cuComplex* output,
    size_t numOutputs, int32_t cudaDevice, cudaStream_t cudaStream) {
  return firEntry(kElemReal, true, decimation, taps, tapCount, input, output, numOutputs, cudaDevice, cudaStream);
}

// 2^64 * frac(f / fs), shared with chain.cu
namespace b200sdr {
uint64_t phaseStepOf(double frequency, double sampleRate) {
  long double cyc = static_cast<long double>(frequency) / static_cast<long double>(sampleRate);
  cyc -= floorl(cyc);
  const long double scaled = cyc * 18446744073709551616.0L;
  if (scaled >= 18446744073709551615.0L) return 0;
  if (scaled >= 9223372036854775808.0L) return static_cast<uint64_t>(llroundl(scaled - 18446744073709551616.0L));
  return static_cast<uint64_t>(llroundl(scaled));
}
}  // namespace b200sdr

GSDR_EXPORT cudaError_t gsdrFmDemod(
    float rfSampleRate, float tunedFrequency, float channelFrequency, float channelFmDeviation,
    size_t rfLowPassDecimation, size_t firstSampleOffset, const float* lowPassTaps, size_t lowPassTapCount,
    const cuComplex* input, float* output, size_t outputCount, int32_t cudaDevice, cudaStream_t cudaStream) {
  DeviceGuard guard(cudaDevice);
  if (guard.status != cudaSuccess) return guard.status;
  if (outputCount == 0) return cudaSuccess;
  if (lowPassTapCount == 0 || lowPassTapCount > 0xffffffffull || rfLowPassDecimation > 0xffffffffull) return cudaErrorInvalidValue;
  FirParams prm {};
  prm.in = input;
  prm.out = output;
  prm.taps = lowPassTaps;
  prm.nOut = outputCount;
  prm.T = static_cast<unsigned>(lowPassTapCount);
  prm.D = rfLowPassDecimation == 0 ? 1u : static_cast<unsigned>(rfLowPassDecimation);
  prm.nIn = outputCount * static_cast<unsigned long long>(prm.D) + lowPassTapCount;
  prm.firstIndex = firstSampleOffset;
  prm.phaseStep = phaseStepOf(static_cast<double>(tunedFrequency) - static_cast<double>(channelFrequency), rfSampleRate);
  prm.mod = kModFm;
  prm.gain = (rfSampleRate / static_cast<float>(prm.D)) / (2.0f * 3.14159265358979323846f * channelFmDeviation);
  prm.inScale = 1.0f;
  return launchFir(kElemComplex, false, true, prm, cudaStream);
}
