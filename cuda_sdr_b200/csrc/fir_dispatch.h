// Host-side routing of one decimating-FIR launch (rows fast path vs direct fallback).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200sdr {

struct FirParams;

struct FirRoute {
  bool rows;      // true: rowsKernel; false: directKernel
  unsigned M;     // ceil(T / D)
  unsigned MP;    // partial sums kept per row (= M, 1..8)
  unsigned TS;    // row stride of the transposed tap table (next power of two of MP)
  int rptIdx;
  unsigned rpt, rowsPerTile, outPerTile, smemBytes;
};

FirRoute planFir(int elem, bool tapsComplex, const void* in, unsigned T, unsigned D, int mod);
cudaError_t launchFir(int elem, bool tapsComplex, bool mix, FirParams prm, cudaStream_t stream);
// `batch` independent streams with the given element strides through the direct kernel (one launch, grid.y = batch)
cudaError_t launchFirBatched(int elem, FirParams prm, unsigned batch, cudaStream_t stream);
const char* firVariantName(int elem, bool tapsComplex, bool mix, const FirRoute& route, char* buf, size_t bufLen);
uint64_t phaseStepOf(double frequency, double sampleRate);
// register-tiled kernel for many taps per kept output (fir_window.cu)
bool windowEligible(int elem, bool tapsComplex, bool mix, const FirParams& prm);
cudaError_t launchWindow(int elem, FirParams prm, cudaStream_t stream);
cudaError_t launchWindowBatched(int elem, FirParams prm, unsigned batch, cudaStream_t stream);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel, size): the FIR entry points are called per block
// of a stream and the attribute call is a driver round trip
cudaError_t ensureDynamicSmem(const void* kernel, int bytes);

// rows per thread of the high-RPT variant for MP partial sums (register budget)
constexpr unsigned rowsRptHigh(unsigned MP) { return MP <= 4 ? 4u : 2u; }

using FirKernel = void (*)(const FirParams);
extern const FirKernel kRowsInt8Mix[16], kRowsInt8Plain[16], kRowsCf32Mix[16], kRowsCf32Plain[16], kRowsCf32PlainWide[3];

}  // namespace b200sdr
