// Planning and launch of the fused persistent chain kernel.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "chain_dispatch.h"
#include "chain_kernels.cuh"

namespace b200sdr {

namespace {

int envInt(const char* name, int fallback) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : fallback;
}

ChainKernel chainKernelFor(int elem, bool mix, unsigned MP, unsigned rptIdx, unsigned conv) {
  const ChainKernel* table = elem == kElemInt8Complex ? (mix ? kChainInt8Mix : kChainInt8Plain) : (mix ? kChainCf32Mix : kChainCf32Plain);
  return table[(MP - 1) * 6 + rptIdx * 2 + conv];
}

constexpr unsigned kSmemPerSm = 227u * 1024u;   // usable shared memory per SM (and per CTA) on B200
constexpr unsigned kSmemPerCtaReserve = 1024u;  // the runtime reserves 1 KB per resident CTA

}  // namespace

ChainPlan planChain(int elem, bool mix, const void* in, unsigned T1, unsigned D1, int mod, unsigned T2, unsigned D2, int device) {
  ChainPlan p {};
  p.fused = false;
  // B200SDR_FUSED: 0 = always the two-kernel path, 1 = fused wherever the shape allows, unset = fused where it is the
  // measured win (int8 input through the tensor cores with <= 4 taps per decimation phase)
  const int fusedMode = envInt("B200SDR_FUSED", -1);
  if (fusedMode == 0) return p;
  if (elem != kElemInt8Complex && elem != kElemComplex) return p;
  if (mod != kModAm && mod != kModFm) return p;
  if (T1 == 0 || D1 == 0 || T2 == 0 || D2 == 0) return p;
  const unsigned es = elem == kElemInt8Complex ? 2u : 8u;
  const unsigned vec = 16u / es;
  p.M = (T1 + D1 - 1) / D1;
  if (D1 % vec != 0 || (reinterpret_cast<uintptr_t>(in) & 15u) != 0 || p.M > 8) return p;
  p.MP = p.M;
  p.TS = static_cast<unsigned>(tapStride(static_cast<int>(p.MP)));
  const unsigned fm = mod == kModFm ? 1u : 0u;
  // int8 input goes through the int8 tensor cores unless B200SDR_CHAIN_MMA=0 (ablation: packed-FP32 CUDA-core route)
  p.conv = (elem == kElemInt8Complex && envInt("B200SDR_CHAIN_MMA", 1) != 0) ? 1u : 0u;
  p.kSteps = (2u * D1 + 31u) / 32u;
  p.bFragWords = p.conv ? 3u * ((2u * p.MP + 7u) / 8u) * p.kSteps * 64u : 0u;
  if (fusedMode < 0 && !(p.conv && p.MP <= 4)) return p;
  int forcedRpt = envInt("B200SDR_CHAIN_RPT", 0);
  const int forcedStages = envInt("B200SDR_CHAIN_STAGES", 0);
  const int forcedCtas = envInt("B200SDR_CHAIN_CTAS", 0);

  int sms = kSmCount;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);

  // Candidates: (rows per thread, compute warps per CTA, ring depth).  The main loop is bound by instruction dispatch
  // and loses only ~4 % at 2 resident warps per scheduler (tools/loop_bench.cu), so the first choice is the big tile
  // (4 rows per thread) with a 2-deep ring; smaller shapes follow for long rows (large D1) that do not fit.
  const int forcedWarps = envInt("B200SDR_CHAIN_WARPS", 0);
  // tensor route: many small row blocks (8 warps x 2 rows per lane) measured best -- the per-tile work is a short
  // latency chain, so it wants warps, not rows per lane; CUDA-core route: the big tile amortises the table reads
  const unsigned big = p.MP <= 4 ? 4u : 2u;
  const unsigned rptChoices[3] = {p.conv ? 2u : big, p.conv ? big : 2u, 1u};
  if (forcedRpt != static_cast<int>(big) && forcedRpt != 2 && forcedRpt != 1) forcedRpt = 0;  // not available for this MP
  for (unsigned stages : {2u, 3u, 4u, 1u}) {
    if (forcedStages && static_cast<unsigned>(forcedStages) != stages) continue;
    for (unsigned rpt : rptChoices) {
      if (forcedRpt && static_cast<unsigned>(forcedRpt) != rpt) continue;
      const unsigned warpOrderMma[6] = {8u, 6u, 4u, 12u, 16u, 2u}, warpOrderFp32[6] = {4u, 8u, 6u, 12u, 16u, 2u};
      for (unsigned wi = 0; wi < 6; wi++) {
        const unsigned warps = p.conv ? warpOrderMma[wi] : warpOrderFp32[wi];
        if (forcedWarps && static_cast<unsigned>(forcedWarps) != warps) continue;
        const int forcedAudio = envInt("B200SDR_CHAIN_AUDIO_WARPS", 0);
        // the audio FIR costs T2/D2 multiply-adds per demodulated sample: one warp keeps up with C2's 13, WBFM's 55 needs four
        unsigned audioWarps = (T2 / D2 + 15u) / 16u;
        if (audioWarps < 1u) audioWarps = 1u;
        if (audioWarps > 4u) audioWarps = 4u;
        if (forcedAudio >= 1 && forcedAudio <= 4) audioWarps = static_cast<unsigned>(forcedAudio);
        if (warps + audioWarps > (rpt == 4 ? 6u : rpt == 2 ? 10u : 18u)) continue;  // __launch_bounds__ of the kernel (register budget)
        if (32u * rpt <= p.M - 1 + fm) continue;
        const unsigned outPerWarp = 32u * rpt - (p.M - 1) - fm;
        const unsigned outPerTile = warps * outPerWarp;
        const unsigned dmCapacity = (outPerTile + T2 + 3u) & ~3u;
        const ChainSmem2 lay = chainSmemLayout2(D1, p.TS, p.M, rpt, warps, es, fm != 0, T2, dmCapacity, stages, p.conv != 0, p.bFragWords);
        if (lay.total > kSmemPerSm - kSmemPerCtaReserve) continue;
        unsigned ctas = kSmemPerSm / (lay.total + kSmemPerCtaReserve);
        const unsigned maxByThreads = 2048u / (32u * (warps + audioWarps));
        if (ctas > maxByThreads) ctas = maxByThreads;
        {  // registers: ask the runtime what actually fits (needs a device; without one the shared-memory bound stands)
          const ChainKernel k = chainKernelFor(elem, mix, p.MP, rpt == 4 ? 2u : rpt == 2 ? 1u : 0u, p.conv);
          int byOccupancy = 0;
          if (cudaFuncSetAttribute(reinterpret_cast<const void*>(k), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerSm)) == cudaSuccess &&
              cudaOccupancyMaxActiveBlocksPerMultiprocessor(&byOccupancy, k, static_cast<int>(32u * (warps + audioWarps)), lay.total) == cudaSuccess &&
              byOccupancy > 0 && static_cast<unsigned>(byOccupancy) < ctas) {
            ctas = static_cast<unsigned>(byOccupancy);
          }
          (void)cudaGetLastError();
        }
        if (ctas > 4) ctas = 4;
        if (forcedCtas) ctas = ctas < static_cast<unsigned>(forcedCtas) ? ctas : static_cast<unsigned>(forcedCtas);
        if (ctas * warps < 8 && !(forcedStages || forcedWarps || forcedRpt || forcedCtas)) continue;  // want >= 2 compute warps per scheduler
        p.fused = true;
        p.rpt = rpt;
        p.rptIdx = rpt == 4 ? 2u : rpt == 2 ? 1u : 0u;
        p.computeWarps = warps;
        p.audioWarps = audioWarps;
        p.tileRows = warps * outPerWarp + (p.M - 1) + fm;
        p.outPerTile = outPerTile;
        p.stages = stages;
        p.dmCapacity = dmCapacity;
        p.smemBytes = lay.total;
        p.ctasPerSm = ctas;
        p.grid = static_cast<unsigned>(sms) * ctas;
        return p;
      }
    }
  }
  return p;
}

cudaError_t launchChain(int elem, bool mix, const ChainPlan& plan, ChainParams prm, cudaStream_t stream) {
  if (!plan.fused) return cudaErrorInvalidConfiguration;
  if (prm.nAudio == 0) return cudaSuccess;
  prm.M = plan.M;
  prm.stages = plan.stages;
  prm.audioWarps = plan.audioWarps;
  prm.dmCapacity = plan.dmCapacity;
  const ChainKernel k = chainKernelFor(elem, mix, plan.MP, plan.rptIdx, plan.conv);
  if (plan.smemBytes > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(k), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(kSmemPerSm));
    if (e != cudaSuccess) return e;
  }
  unsigned grid = plan.grid;
  if (static_cast<unsigned long long>(grid) > prm.nAudio) grid = static_cast<unsigned>(prm.nAudio);
  // tools only: B200SDR_CHAIN_PROFILE=1 prints where the compute warps spend their cycles (synchronises!)
#ifdef B200SDR_CHAIN_PROFILE_BUILD
  static const bool profile = envInt("B200SDR_CHAIN_PROFILE", 0) != 0;
#else
  constexpr bool profile = false;  // build with NVCC_EXTRA=-DB200SDR_CHAIN_PROFILE_BUILD to enable (tools/chain_profile.sh)
#endif
  unsigned long long* prof = nullptr;
  if (profile) {
    cudaMalloc(&prof, sizeof(unsigned long long) * 6 * grid * plan.computeWarps);
    cudaMemset(prof, 0, sizeof(unsigned long long) * 6 * grid * plan.computeWarps);
    prm.prof = prof;
  }
  k<<<grid, 32u * (plan.computeWarps + plan.audioWarps), plan.smemBytes, stream>>>(prm);
  if (profile) {
    cudaStreamSynchronize(stream);
    std::vector<unsigned long long> h(6ull * grid * plan.computeWarps);
    cudaMemcpy(h.data(), prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(prof);
    if (const char* dump = std::getenv("B200SDR_CHAIN_PROFILE_DUMP")) {  // per (CTA, warp): sm, wait-tile, main, rest, wait-line, total
      if (FILE* f = fopen(dump, "w")) {
        for (size_t i = 0; i + 5 < h.size(); i += 6)
          fprintf(f, "%zu %llu %llu %llu %llu %llu %llu\n", i / 6 / plan.computeWarps, h[i + 4] >> 40, h[i], h[i + 1], h[i + 2], h[i + 3], h[i + 5]);
        fclose(f);
      }
    }
    for (size_t i = 4; i < h.size(); i += 6) h[i] &= (1ull << 40) - 1;
    double sum[6] = {0, 0, 0, 0, 0, 0}, longest = 0;
    for (size_t i = 0; i < h.size(); i++) {
      sum[i % 6] += static_cast<double>(h[i]);
      if (i % 6 == 5 && static_cast<double>(h[i]) > longest) longest = static_cast<double>(h[i]);
    }
    const double n = static_cast<double>(h.size() / 6);
    fprintf(stderr,
            "[chain profile] per compute warp, cycles: wait-tile %.0f, main loop %.0f, exchange+demod %.0f, wait-line %.0f, prologue %.0f, "
            "whole warp %.0f (longest %.0f)\n",
            sum[0] / n, sum[1] / n, sum[2] / n, sum[3] / n, sum[4] / n, sum[5] / n, longest);
  }
  return launchStatus();
}

const char* chainVariantName(int elem, bool mix, const ChainPlan& plan, char* buf, size_t bufLen) {
  const char* e = elem == kElemInt8Complex ? "int8c" : "cf32";
  snprintf(buf, bufLen, "chain<%s,mix=%d,MP=%u,RPT=%u,conv=%s>(warps=%u+%u,tileRows=%u,stages=%u,smem=%u,ctas/SM=%u,grid=%u)", e,
           mix ? 1 : 0, plan.MP, plan.rpt, plan.conv ? "imma" : "fp32x2", plan.computeWarps, plan.audioWarps, plan.tileRows, plan.stages, plan.smemBytes,
           plan.ctasPerSm, plan.grid);
  return buf;
}

}  // namespace b200sdr
