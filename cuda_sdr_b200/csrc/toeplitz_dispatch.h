// Host-side planning and launch of toepKernel (toeplitz_kernels.cuh): the int8 chain whose RF stage is one int8 GEMM over
// a Toeplitz view of the raw input.
#pragma once

#include <cuda.h>  // CUtensorMap and the encoder's signature (types only: no link-time dependency on libcuda)
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace b200sdr {

struct ToepParams;

// cuTensorMapEncodeTiled, fetched from the driver at run time (nullptr without a driver)
using EncodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiled encodeTiled();

struct ToepPlan {
  bool ok;       // false: shape not supported (D1 not a multiple of 8, no audio FIR, tables too large, or disabled by B200SDR_TOEPLITZ=0)
  bool magic;    // accumulators start at the float magic number (set by buildToeplitzFragments from the digits' worst-case sums)
  unsigned G;    // m-tiles (64 outputs each) per warp block
  unsigned Q, KS;
  unsigned NW, NA, S;
  unsigned blockBytes, boxBytes, slotBytes, OTW, OT, span;
  unsigned smemBytes, ctasPerSm, grid;
};

ToepPlan planToeplitz(unsigned T1, unsigned D1, int mod, unsigned T2, unsigned D2, int device);
void buildToeplitzFragments(const float* taps, unsigned T1, unsigned D1, bool mix, uint64_t phaseStep, double inScale, ToepPlan& plan,
                            std::vector<uint32_t>& frag, float digitScale[3]);
cudaError_t launchToeplitz(const ToepPlan& plan, ToepParams prm, cudaStream_t stream);
const char* toeplitzVariantName(const ToepPlan& plan, char* buf, size_t bufLen);

}  // namespace b200sdr
