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
  unsigned rpt, rptIdx;
  unsigned conv;                    // 0: CUDA cores (packed FP32, ALU-pipe int8 conversion); 1: int8 tensor cores (int8 input only)
  unsigned kSteps, bFragWords;      // tensor route: ceil(2*D1/32) k-steps; words of the B-fragment table
  unsigned computeWarps, audioWarps;  // per CTA; blockDim = 32 * (computeWarps + audioWarps)
  unsigned tileRows, outPerTile;    // rows staged per tile, demodulated samples it yields
  unsigned stages, dmCapacity;
  unsigned smemBytes, ctasPerSm, grid;
};

// elem: kElemInt8Complex / kElemComplex.  `in` is only inspected for its 16-byte alignment.
ChainPlan planChain(int elem, bool mix, const void* in, unsigned T1, unsigned D1, int mod, unsigned T2, unsigned D2, int device);
cudaError_t launchChain(int elem, bool mix, const ChainPlan& plan, ChainParams prm, cudaStream_t stream);
const char* chainVariantName(int elem, bool mix, const ChainPlan& plan, char* buf, size_t bufLen);

using ChainKernel = void (*)(const ChainParams);
// [MP-1][rptIdx (0: 1 row/lane, 1: 2 rows/lane, 2: 4 rows/lane; MP > 4: 2)][conv (0: CUDA cores, 1: tensor cores)]
extern const ChainKernel kChainInt8Mix[48], kChainInt8Plain[48], kChainCf32Mix[48], kChainCf32Plain[48];

}  // namespace b200sdr
