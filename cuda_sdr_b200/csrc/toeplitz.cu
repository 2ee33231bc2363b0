// Planning, constant tables and launch of toepKernel (toeplitz_kernels.cuh).
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "digits.h"
#include "toeplitz_dispatch.h"
#include "toeplitz_kernels.cuh"

namespace b200sdr {

// cuTensorMapEncodeTiled lives in the driver library; fetch it through the runtime so libb200sdr.so has no link-time
// dependency on libcuda (the library must load, and export its symbols, on a box without a driver)
EncodeTiled encodeTiled() {
  static EncodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiled>(p);
  }();
  return fn;
}

namespace {

int envInt(const char* name, int fallback) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : fallback;
}

constexpr unsigned kSmemPerSm = 227u * 1024u;
constexpr unsigned kSmemPerCtaReserve = 1024u;

using ToepKernel = void (*)(const ToepParams, const CUtensorMap);

unsigned swizzleSpan(unsigned D1) { return (8u * D1) % 128u == 64u ? 64u : 128u; }
ToepKernel toepKernelFor(unsigned G, bool magic) {
  if (G == 2) return magic ? toepKernel<2, true> : toepKernel<2, false>;
  return magic ? toepKernel<1, true> : toepKernel<1, false>;
}

}  // namespace

ToepPlan planToeplitz(unsigned T1, unsigned D1, int mod, unsigned T2, unsigned D2, int device) {
  ToepPlan p {};
  p.ok = false;
  // B200SDR_TOEPLITZ=0: fall back to chainKernel / the two-kernel path; B200SDR_FUSED=0: two-kernel path everywhere
  if (envInt("B200SDR_TOEPLITZ", 1) == 0 || envInt("B200SDR_FUSED", -1) == 0) return p;
  if (mod != kModAm && mod != kModFm) return p;
  if (T1 == 0 || D1 == 0 || T2 == 0 || D2 == 0) return p;
  if (D1 % 8u != 0) return p;  // rows must start on 16-byte boundaries (TMA bulk copies, 128-bit fragment loads)
  const unsigned fm = mod == kModFm ? 1u : 0u;
  const unsigned kBytes = 2u * T1 + 6u * D1;  // bytes of an A-row that carry a non-zero B row
  p.KS = (kBytes + 31u) / 32u;
  p.Q = (p.KS + 1u) / 2u;
  if (p.Q * 1536u > 96u * 1024u) return p;  // B fragments must leave room for the rings
  const unsigned AS = 8u * D1;

  int sms = kSmCount;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  (void)cudaGetLastError();

  // audio warps take tiles in turn; the FIR costs T2/D2 multiply-adds per demodulated sample (C2: 13, WBFM: 55)
  unsigned audioWarps = T2 / D2 >= 32u ? 4u : 2u;
  const int forcedAudio = envInt("B200SDR_TOEP_AUDIO_WARPS", 0);
  if (forcedAudio >= 1 && forcedAudio <= 4) audioWarps = static_cast<unsigned>(forcedAudio);
  const int forcedG = envInt("B200SDR_TOEP_G", 0), forcedWarps = envInt("B200SDR_TOEP_WARPS", 0);
  const int forcedStages = envInt("B200SDR_TOEP_STAGES", 0), forcedCtas = envInt("B200SDR_TOEP_CTAS", 0);

  // Candidates in order of preference; the first one that puts >= 8 compute warps on an SM wins, else the one with the most.
  unsigned bestWarps = 0;
  for (unsigned G : {2u, 1u}) {
    if (forcedG && static_cast<unsigned>(forcedG) != G) continue;
    for (unsigned stages : {2u, 3u}) {
      if (forcedStages && static_cast<unsigned>(forcedStages) != stages) continue;
      for (unsigned warps : {4u, 8u, 6u, 7u, 5u, 3u, 2u}) {
        if (forcedWarps && static_cast<unsigned>(forcedWarps) != warps) continue;
        if (warps + audioWarps > 12u || warps * stages > 48u) continue;  // __launch_bounds__(384), barrier table
        const unsigned W = swizzleSpan(D1);
        const unsigned blockBytes = (16u * G - 1u) * AS + 64u * p.Q;
        const unsigned boxBytes = (blockBytes + W - 1u) / W * W;
        if (boxBytes / W > 256u) continue;  // box extent limit of a tensor map
        const unsigned atom = W == 64u ? 512u : 1024u;  // the swizzle pattern repeats every 4 / 8 rows of 128 bytes
        const unsigned slotBytes = (boxBytes + atom - 1u) / atom * atom;
        const unsigned OTW = 64u * G - fm, OT = warps * OTW;
        const unsigned span = (T2 - 1u + OT - 1u) / OT;  // tiles an audio window reaches back
        if (span > kToepLines - 2u) continue;
        if (((T2 + 3u) & ~3u) > OT) continue;  // keep the mirror inside the first tile of the ring (longer audio filters: other routes)
        const ToepSmem lay = toepSmemLayout(p.Q, T2, OT, warps, stages, slotBytes);
        if (lay.total > kSmemPerSm - kSmemPerCtaReserve) continue;
        unsigned ctas = kSmemPerSm / (lay.total + kSmemPerCtaReserve);
        const unsigned maxByThreads = 2048u / (32u * (warps + audioWarps));
        if (ctas > maxByThreads) ctas = maxByThreads;
        {
          const ToepKernel k = toepKernelFor(G, true);
          int byOccupancy = 0;
          if (cudaFuncSetAttribute(reinterpret_cast<const void*>(k), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerSm)) == cudaSuccess &&
              cudaOccupancyMaxActiveBlocksPerMultiprocessor(&byOccupancy, k, static_cast<int>(32u * (warps + audioWarps)), lay.total) == cudaSuccess &&
              byOccupancy > 0 && static_cast<unsigned>(byOccupancy) < ctas) {
            ctas = static_cast<unsigned>(byOccupancy);
          }
          (void)cudaGetLastError();
        }
        if (ctas > 4u) ctas = 4u;
        if (forcedCtas && static_cast<unsigned>(forcedCtas) < ctas) ctas = static_cast<unsigned>(forcedCtas);
        const unsigned total = ctas * warps;
        const bool forced = forcedG || forcedWarps || forcedStages || forcedCtas;
        if (total > bestWarps || (forced && !p.ok)) {
          bestWarps = total;
          p.ok = true;
          p.G = G;
          p.NW = warps;
          p.NA = audioWarps;
          p.S = stages;
          p.blockBytes = blockBytes;
          p.boxBytes = boxBytes;
          p.slotBytes = slotBytes;
          p.OTW = OTW;
          p.OT = OT;
          p.span = span;
          p.smemBytes = lay.total;
          p.ctasPerSm = ctas;
          p.grid = static_cast<unsigned>(sms) * ctas;
          // B200SDR_TOEP_GRID: fewer CTAs than the machine holds, so that each one walks many tiles (parity tests use it to
          // wrap the demod ring several times on small inputs)
          const int forcedGrid = envInt("B200SDR_TOEP_GRID", 0);
          if (forcedGrid > 0 && static_cast<unsigned>(forcedGrid) < p.grid) p.grid = static_cast<unsigned>(forcedGrid);
        }
        if (total >= 8u) return p;
      }
    }
    // two m-tiles per warp halve the B-fragment traffic per output: keep G = 2 whenever it puts >= 4 warps on an SM
    if (G == 2u && p.ok && bestWarps >= 4u) return p;
  }
  return p;
}

// B (see toeplitz_kernels.cuh) quantised to three signed int8 digits and laid out in fragment order.
void buildToeplitzFragments(const float* taps, unsigned T1, unsigned D1, bool mix, uint64_t phaseStep, double inScale, ToepPlan& plan,
                            std::vector<uint32_t>& frag, float digitScale[3]) {
  const unsigned rows = plan.Q * 64u;  // physical bytes of an A-row covered by the fragment loads
  std::vector<double> B(static_cast<size_t>(rows) * 8u, 0.0);
  double bMax = 0.0;
  for (unsigned j = 0; j < T1; j++) {
    double cr = static_cast<double>(taps[j]) * inScale, ci = 0.0;
    if (mix) {
      const double frac = static_cast<double>(static_cast<int64_t>(phaseStep * j)) * (1.0 / 18446744073709551616.0);
      const double phi = 6.283185307179586476925286766559 * frac;
      ci = cr * std::sin(phi);
      cr = cr * std::cos(phi);
    }
    bMax = std::fmax(bMax, std::fmax(std::fabs(cr), std::fabs(ci)));
    for (unsigned n = 0; n < 4; n++) {
      const size_t byte = 2u * (static_cast<size_t>(D1) * n + j);
      B[byte * 8u + 2u * n] = cr;
      B[(byte + 1u) * 8u + 2u * n] = -ci;
      B[byte * 8u + 2u * n + 1u] = ci;
      B[(byte + 1u) * 8u + 2u * n + 1u] = cr;
    }
  }
  const FixedPoint24 fx = fixedPoint24For(bMax);
  for (int d = 0; d < 3; d++) digitScale[d] = fx.digitScale[d];
  frag.assign(static_cast<size_t>(plan.Q) * 32u * 12u, 0u);
  double colAbs[3][8] = {};
  for (unsigned q = 0; q < plan.Q; q++)
    for (unsigned lane = 0; lane < 32; lane++) {
      const unsigned g = lane >> 2, t = lane & 3u;
      for (unsigned ksub = 0; ksub < 2; ksub++)
        for (unsigned half = 0; half < 2; half++) {
          unsigned word[3] = {0, 0, 0};
          for (unsigned e = 0; e < 4; e++) {
            const size_t byte = 32u * (2u * q + ksub) + 16u * half + 4u * t + e;
            int dg[3];
            balancedDigits(B[byte * 8u + g], fx.scale, dg);
            for (int d = 0; d < 3; d++) {
              word[d] |= (static_cast<unsigned>(dg[d]) & 0xffu) << (8u * e);
              colAbs[d][g] += std::abs(dg[d]);
            }
          }
          for (unsigned d = 0; d < 3; d++) frag[(static_cast<size_t>(q) * 32u + lane) * 12u + (ksub * 3u + d) * 2u + half] = word[d];
        }
    }
  // the accumulators may start at the float magic number (their bits are then the float) iff |sum| < 2^22 for any int8 input
  double worst = 0.0;
  for (int d = 0; d < 3; d++)
    for (int n = 0; n < 8; n++) worst = std::fmax(worst, colAbs[d][n] * 128.0);
  plan.magic = worst < 4194304.0 && envInt("B200SDR_TOEP_MAGIC", 1) != 0;
}

cudaError_t launchToeplitz(const ToepPlan& plan, ToepParams prm, cudaStream_t stream) {
  if (!plan.ok) return cudaErrorInvalidConfiguration;
  if (prm.nAudio == 0) return cudaSuccess;
  prm.Q = plan.Q;
  prm.KS = plan.KS;
  prm.NW = plan.NW;
  prm.NA = plan.NA;
  prm.S = plan.S;
  prm.slotBytes = plan.slotBytes;
  prm.blockBytes = plan.blockBytes;
  prm.span = plan.span;
  static const int prefetch = envInt("B200SDR_TOEP_PREFETCH", 0);
  prm.prefetch = prefetch < 0 ? 0u : static_cast<unsigned>(prefetch);
  prm.boxBytes = plan.boxBytes;
  // the input as a tensor {W bytes, chunk (stride W), shift (stride 16 B)}: coordinate (0, s / W, (s % W) / 16) addresses any
  // 16-byte-aligned offset s; chunks are limited so that no shift reads past the end (the kernel patches the last bytes)
  const unsigned W = swizzleSpan(prm.D1);
  prm.wShift = W == 64u ? 6u : 7u;
  prm.swzMask = W == 64u ? 0x30u : 0x70u;
  if (prm.nInBytes < 2ull * W) return cudaErrorInvalidValue;
  const unsigned long long chunks = (prm.nInBytes - (W - 16u)) / W;
  prm.tmaEnd = chunks * W;
  const EncodeTiled encode = encodeTiled();
  if (!encode) return cudaErrorNotSupported;
  CUtensorMap tmap;
  static const int l2promo = envInt("B200SDR_TOEP_L2PROMO", static_cast<int>(CU_TENSOR_MAP_L2_PROMOTION_L2_128B));  // 0 none, 1 64B, 2 128B, 3 256B
  const cuuint64_t gdim[3] = {W, chunks, W / 16u};
  const cuuint64_t gstride[2] = {W, 16u};
  const cuuint32_t box[3] = {W, plan.boxBytes / W, 1u};
  const cuuint32_t estride[3] = {1u, 1u, 1u};
  if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<unsigned char*>(prm.in), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
             W == 64u ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, static_cast<CUtensorMapL2promotion>(l2promo),
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  const ToepKernel k = toepKernelFor(plan.G, plan.magic);
  if (plan.smemBytes > 48u * 1024u) {
    const cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(k), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerSm));
    if (e != cudaSuccess) return e;
  }
  unsigned grid = plan.grid;
  if (static_cast<unsigned long long>(grid) > prm.nAudio) grid = static_cast<unsigned>(prm.nAudio);
  // programmatic dependent launch: this kernel's prologue overlaps the tail of the previous kernel in the stream
  static const bool pdl = envInt("B200SDR_TOEP_PDL", 1) != 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32u * (plan.NW + plan.NA));
  cfg.dynamicSmemBytes = plan.smemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1u : 0u;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, k, prm, tmap);
  if (e != cudaSuccess) return e;
  return launchStatus();
}

const char* toeplitzVariantName(const ToepPlan& plan, char* buf, size_t bufLen) {
  snprintf(buf, bufLen, "toeplitz<int8c,G=%u,%s>(kSteps=%u,warps=%u+%u,block=%u B,stages=%u,smem=%u,ctas/SM=%u,grid=%u)", plan.G,
           plan.magic ? "magic" : "i2f", plan.KS, plan.NW, plan.NA, plan.blockBytes, plan.S, plan.smemBytes, plan.ctasPerSm, plan.grid);
  return buf;
}

}  // namespace b200sdr
