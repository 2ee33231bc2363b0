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
  unsigned MP;    // padded partial-sum count (1, 2, 4, 8)
  int mpIdx, rptIdx;
  unsigned rpt, rowsPerTile, outPerTile, smemBytes;
};

FirRoute planFir(int elem, bool tapsComplex, const void* in, unsigned T, unsigned D, int mod);
cudaError_t launchFir(int elem, bool tapsComplex, bool mix, FirParams prm, cudaStream_t stream);
const char* firVariantName(int elem, bool tapsComplex, bool mix, const FirRoute& route, char* buf, size_t bufLen);
uint64_t phaseStepOf(double frequency, double sampleRate);

}  // namespace b200sdr
