// Host-side planning and launch of the fused persistent chain kernel (chain_kernels.cuh).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200sdr {

struct ChainParams;

struct ChainPlan {
  bool fused;  // false: this shape/alignment has to take the two-kernel path (rowsKernel/directKernel + audio FIR)
  unsigned M, MP, TS;
  unsigned rpt, rptIdx, conv;
  unsigned rowsPerTile, outPerTile;
  unsigned stages, audioParts, dmCapacity;
  unsigned smemBytes, ctasPerSm, grid;
};

// elem: kElemInt8Complex / kElemComplex.  `in` is only inspected for its 16-byte alignment.
ChainPlan planChain(int elem, bool mix, const void* in, unsigned T1, unsigned D1, int mod, unsigned T2, unsigned D2, int device);
cudaError_t launchChain(int elem, bool mix, const ChainPlan& plan, ChainParams prm, cudaStream_t stream);
const char* chainVariantName(int elem, bool mix, const ChainPlan& plan, char* buf, size_t bufLen);

using ChainKernel = void (*)(const ChainParams);
// [MP-1][rptIdx (0: 2 rows/thread, 1: 4 rows/thread; MP > 4 always 2)][conv (0: magic-number/FADD2, 1: sign-extend/I2FP)]
extern const ChainKernel kChainInt8Mix[32], kChainInt8Plain[32], kChainCf32Mix[32], kChainCf32Plain[32];

}  // namespace b200sdr
