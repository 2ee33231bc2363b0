// Planning and launch of the fused persistent chain kernel.
#include <cstdio>
#include <cstdlib>

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
  return table[(MP - 1) * 4 + rptIdx * 2 + conv];
}

constexpr unsigned kSmemPerSm = 227u * 1024u;   // usable shared memory per SM (and per CTA) on B200
constexpr unsigned kSmemPerCtaReserve = 1024u;  // the runtime reserves 1 KB per resident CTA

}  // namespace

ChainPlan planChain(int elem, bool mix, const void* in, unsigned T1, unsigned D1, int mod, unsigned T2, unsigned D2, int device) {
  ChainPlan p {};
  p.fused = false;
  if (envInt("B200SDR_FUSED", 0) == 0) return p;  // opt-in until the warp-specialised version lands
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
  p.conv = static_cast<unsigned>(envInt("B200SDR_CHAIN_CONV", 1)) & 1u;
  int forcedRpt = envInt("B200SDR_CHAIN_RPT", 0);
  const int forcedStages = envInt("B200SDR_CHAIN_STAGES", 0);
  const int forcedCtas = envInt("B200SDR_CHAIN_CTAS", 0);

  int sms = kSmCount;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);

  // Candidates: (rows per thread, ring depth).  Prefer >= 3 resident CTAs per SM so that the barrier-separated
  // exchange / audio phases of one CTA hide behind the main loops of the others; then the largest tile.
  const unsigned rptChoices[2] = {p.MP <= 4 ? 4u : 2u, 2u};
  if (forcedRpt != static_cast<int>(rptChoices[0]) && forcedRpt != 2) forcedRpt = 0;  // not available for this MP
  for (unsigned wantCtas : {3u, 2u, 1u}) {
    for (unsigned rpt : rptChoices) {
      if (forcedRpt && static_cast<unsigned>(forcedRpt) != rpt) continue;
      for (unsigned stages : {2u, 3u, 1u}) {
        if (forcedStages && static_cast<unsigned>(forcedStages) != stages) continue;
        const unsigned rowsPerTile = rpt * kRowsThreads;
        if (rowsPerTile <= p.M - 1 + fm) continue;
        const unsigned outPerTile = rowsPerTile - (p.M - 1) - fm;
        const unsigned dmCapacity = (outPerTile + T2 + 3u) & ~3u;
        const ChainSmem lay = chainSmemLayout(D1, p.TS, p.M, rowsPerTile, es, fm != 0, T2, dmCapacity, stages);
        if (lay.total > kSmemPerSm - kSmemPerCtaReserve) continue;
        unsigned ctas = kSmemPerSm / (lay.total + kSmemPerCtaReserve);
        if (ctas > 6) ctas = 6;
        if (forcedCtas) ctas = ctas < static_cast<unsigned>(forcedCtas) ? ctas : static_cast<unsigned>(forcedCtas);
        if (ctas < wantCtas) continue;
        p.fused = true;
        p.rpt = rpt;
        p.rptIdx = rpt == 4 ? 1u : 0u;
        p.rowsPerTile = rowsPerTile;
        p.outPerTile = outPerTile;
        p.stages = stages;
        p.dmCapacity = dmCapacity;
        p.smemBytes = lay.total;
        p.ctasPerSm = ctas;
        p.grid = static_cast<unsigned>(sms) * ctas;
        // threads per audio dot product: spread the <= nA outputs of a tile over the whole CTA
        const unsigned nAmax = (outPerTile + T2) / D2 + 1;
        unsigned parts = 1;
        while (parts < 8 && nAmax * parts * 2 <= kRowsThreads) parts *= 2;
        const int forcedParts = envInt("B200SDR_CHAIN_PARTS", 0);
        if (forcedParts == 1 || forcedParts == 2 || forcedParts == 4 || forcedParts == 8) parts = static_cast<unsigned>(forcedParts);
        p.audioParts = parts;
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
  prm.rowsPerTile = plan.rowsPerTile;
  prm.outPerTile = plan.outPerTile;
  prm.stages = plan.stages;
  prm.audioParts = plan.audioParts;
  prm.dmCapacity = plan.dmCapacity;
  const ChainKernel k = chainKernelFor(elem, mix, plan.MP, plan.rptIdx, plan.conv);
  if (plan.smemBytes > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(k), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(kSmemPerSm));
    if (e != cudaSuccess) return e;
  }
  unsigned grid = plan.grid;
  if (static_cast<unsigned long long>(grid) > prm.nAudio) grid = static_cast<unsigned>(prm.nAudio);
  k<<<grid, kRowsThreads, plan.smemBytes, stream>>>(prm);
  return launchStatus();
}

const char* chainVariantName(int elem, bool mix, const ChainPlan& plan, char* buf, size_t bufLen) {
  const char* e = elem == kElemInt8Complex ? "int8c" : "cf32";
  snprintf(buf, bufLen, "chain<%s,mix=%d,MP=%u,RPT=%u,conv=%s>(rowsPerTile=%u,stages=%u,audioParts=%u,smem=%u,ctas/SM=%u,grid=%u)", e,
           mix ? 1 : 0, plan.MP, plan.rpt, plan.conv ? "alu" : "magic", plan.rowsPerTile, plan.stages, plan.audioParts, plan.smemBytes,
           plan.ctasPerSm, plan.grid);
  return buf;
}

}  // namespace b200sdr
