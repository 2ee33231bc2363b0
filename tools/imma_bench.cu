// Throughput / latency of the legacy warp-level int8 MMA (mma.sync.m16n8k32.s8 -> SASS IMMA.16832.S8.S8) on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/imma_bench tools/imma_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <stdint.h>

template <int CHAINS>
__global__ void __launch_bounds__(128) k(int* out, int iters, unsigned seed) {
  int acc[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; c++)
    for (int e = 0; e < 4; e++) acc[c][e] = 0;
  unsigned a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(acc[c][0]), "+r"(acc[c][1]), "+r"(acc[c][2]), "+r"(acc[c][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  int s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++)
    for (int e = 0; e < 4; e++) s += acc[c][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
void run(int ctasPerSm) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int* out;
  cudaMalloc(&out, sizeof(int) * sms * ctasPerSm * 128);
  const int iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<CHAINS><<<sms * ctasPerSm, 128>>>(out, iters, 1);
  cudaEventRecord(a);
  k<CHAINS><<<sms * ctasPerSm, 128>>>(out, iters, 1);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double mmas = double(sms) * ctasPerSm * 4 * iters * CHAINS;
  const double cycles = ms * 1e-3 * 1.965e9;
  printf("{\"chains\": %d, \"warps_per_smsp\": %d, \"ms\": %.3f, \"smsp_cycles_per_imma\": %.2f, \"int8_tops\": %.1f}\n", CHAINS, ctasPerSm, ms,
         cycles * sms * 4 / mmas, mmas * 16 * 8 * 32 * 2 / (ms * 1e-3) / 1e12);
  cudaFree(out);
}

int main() {
  run<1>(1);
  run<2>(1);
  run<4>(1);
  run<8>(1);
  run<4>(2);
  run<4>(4);
  run<8>(4);
  return 0;
}
