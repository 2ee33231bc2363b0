// Upper bound of the chain's main loop on B200: tilePartialSums() over a tile that already sits in shared memory
// (no HBM traffic, no barriers), swept over resident warps per SM.  Tells how far the real kernels are from what the
// SM can issue for this instruction mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Icuda_sdr_b200/csrc -Iinclude -o build/loop_bench tools/loop_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "chain_kernels.cuh"

namespace b200sdr {
std::atomic<uint64_t> g_launchCount {0};
}
using namespace b200sdr;

template <int MP, int RPT, int CONV>
__global__ void __launch_bounds__(kRowsThreads) loopKernel(float2* out, unsigned D, int iters, unsigned warpMap) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TS = tapStride(MP);
  float4* W = reinterpret_cast<float4*>(smem);
  float* hT = reinterpret_cast<float*>(smem + D * 16);
  unsigned char* tile = smem + D * 16 + D * TS * 4;
  const unsigned tid = threadIdx.x;
  for (unsigned i = tid; i < D; i += kRowsThreads) W[i] = make_float4(0.01f * i, 0.02f, -0.02f, 0.01f * i);
  for (unsigned i = tid; i < D * TS; i += kRowsThreads) hT[i] = 0.001f * i;
  for (unsigned i = tid; i < RPT * kRowsThreads * D * 2; i += kRowsThreads) tile[i] = static_cast<unsigned char>(i * 7 + blockIdx.x);
  __syncthreads();
  float2 total = make_float2(0.f, 0.f);
  for (int it = 0; it < iters; it++) {
    float2 acc[RPT][MP];
    if (warpMap) {
      tilePartialSums<kElemInt8Complex, true, MP, RPT, CONV>(tile + (tid >> 5) * (32 * RPT - 2) * D * 2, hT, W, D, tid & 31u, 32u, acc);
    } else {
      tilePartialSums<kElemInt8Complex, true, MP, RPT, CONV>(tile, hT, W, D, tid, kRowsThreads, acc);
    }
#pragma unroll
    for (int i = 0; i < RPT; i++)
#pragma unroll
      for (int m = 0; m < MP; m++) {
        total.x += acc[i][m].x;
        total.y += acc[i][m].y;
      }
  }
  out[blockIdx.x * kRowsThreads + tid] = total;
}

template <int MP, int RPT, int CONV>
void run(int ctasPerSm, unsigned D, int iters, unsigned warpMap = 0) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  constexpr int TS = tapStride(MP);
  const unsigned need = D * 16 + D * TS * 4 + RPT * kRowsThreads * D * 2;
  // pad the dynamic shared memory so that exactly ctasPerSm CTAs fit on an SM
  unsigned smemBytes = (227u * 1024u) / ctasPerSm - 1024u;
  if (smemBytes < need) return;
  auto k = loopKernel<MP, RPT, CONV>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  float2* out;
  cudaMalloc(&out, sizeof(float2) * sms * ctasPerSm * kRowsThreads);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<<<sms * ctasPerSm, kRowsThreads, smemBytes>>>(out, D, iters, warpMap);
  cudaEventRecord(a);
  k<<<sms * ctasPerSm, kRowsThreads, smemBytes>>>(out, D, iters, warpMap);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  const double sampleRows = double(sms) * ctasPerSm * kRowsThreads * RPT * D * iters;
  const double clk = 1.965e9;
  const double cyclesPerWarpSampleRow = (ms * 1e-3 * clk) * (sms * 4.0) / (sampleRows / 32.0);
  printf("{\"warpMap\": %u, \"MP\": %d, \"RPT\": %d, \"conv\": %d, \"ctas_per_sm\": %d, \"warps_per_smsp\": %d, \"ms\": %.4f, \"Tsps\": %.3f, "
         "\"smsp_cycles_per_warp_sample_row\": %.2f, \"err\": \"%s\"}\n",
         warpMap, MP, RPT, CONV, ctasPerSm, ctasPerSm, ms, sampleRows / (ms * 1e-3) / 1e12, cyclesPerWarpSampleRow,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  const unsigned D = 40;
  const int iters = 400;
  for (int ctas : {1, 2, 3, 4, 5}) {
    run<3, 4, 1>(ctas, D, iters, 0);
    run<3, 4, 1>(ctas, D, iters, 1);
    run<3, 4, 0>(ctas, D, iters, 1);
    run<3, 2, 1>(ctas, D, iters, 1);
  }
  return 0;
}
