// Pipe-throughput microbenchmarks for B200 (sm_100a): FP32 FFMA / packed FFMA2 / FADD / PRMT / LOP3 / I2FP.
// Used to pick the K1 inner-loop instruction mix and to measure the FP32 roof quoted in DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench tools/microbench.cu && build/microbench
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CHECK(x)                                                                      \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);  \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

constexpr int kChains = 16;
constexpr int kIters = 4096;

enum Op { FFMA_RRR, FFMA_IMM, FFMA2, FADD, FADD2, PRMT, LOP3, I2FP, MIX_K1, MIX_K1_F2 };

template <int OP>
__global__ void __launch_bounds__(256) bench(float* out, float seed, unsigned iseed) {
  float a[kChains], b = seed, c = seed * 0.5f;
  float2 a2[kChains];
  unsigned u[kChains];
#pragma unroll
  for (int i = 0; i < kChains; i++) {
    a[i] = seed + i + threadIdx.x;
    a2[i] = make_float2(a[i], a[i] + 1.0f);
    u[i] = iseed + i * 77u + threadIdx.x;
  }
  const float2 b2 = make_float2(b, c), c2 = make_float2(c, b);
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int i = 0; i < kChains; i++) {
      if constexpr (OP == FFMA_RRR) a[i] = fmaf(a[i], b, c);
      if constexpr (OP == FFMA_IMM) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      if constexpr (OP == FFMA2) a2[i] = __ffma2_rn(a2[i], b2, c2);
      if constexpr (OP == FADD) a[i] = a[i] + b;
      if constexpr (OP == FADD2) a2[i] = __fadd2_rn(a2[i], b2);
      if constexpr (OP == PRMT) u[i] = __byte_perm(u[i], iseed, 0x7250);
      if constexpr (OP == LOP3) u[i] = (u[i] ^ iseed) & 0x7fffffffu | 0x00010000u;
      if constexpr (OP == I2FP) u[i] = __float_as_uint(static_cast<float>(static_cast<int>(u[i])));
      if constexpr (OP == MIX_K1) {  // K1-like mix: per 2 FFMA one PRMT (different pipes)
        a[i] = fmaf(a[i], b, c);
        u[i] = __byte_perm(u[i], iseed, 0x7250);
        a[i] = fmaf(a[i], c, b);
      }
      if constexpr (OP == MIX_K1_F2) {
        a2[i] = __ffma2_rn(a2[i], b2, c2);
        u[i] = __byte_perm(u[i], iseed, 0x7250);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < kChains; i++) s += a[i] + a2[i].x + a2[i].y + __uint_as_float(u[i] & 0x3fffffffu);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, double opsPerIterPerChain, double flopsPerOp, int sms, double clockGHz, float* out) {
  const int blocks = sms * 8, threads = 256;
  bench<OP><<<blocks, threads>>>(out, 1.0f, 12345u);
  CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CHECK(cudaEventRecord(e0));
    bench<OP><<<blocks, threads>>>(out, 1.0f, 12345u);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double laneOps = double(blocks) * threads * kIters * kChains * opsPerIterPerChain;
  const double perSec = laneOps / (best * 1e-3);
  printf("{\"op\": \"%s\", \"ms\": %.4f, \"lane_ops_per_s\": %.4e, \"lane_ops_per_clk_per_sm\": %.2f, \"tflops\": %.2f}\n", name, best,
         perSec, perSec / sms / (clockGHz * 1e9), perSec * flopsPerOp / 1e12);
}

int main() {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  int clockKHz = 0;
  CHECK(cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, 0));
  const double ghz = clockKHz / 1e6;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz_max\": %.3f, \"note\": \"per-clk figures assume the max clock\"}\n", prop.name,
         prop.multiProcessorCount, ghz);
  float* out;
  CHECK(cudaMalloc(&out, sizeof(float) * prop.multiProcessorCount * 8 * 256));
  const int sms = prop.multiProcessorCount;
  run<FFMA_RRR>("ffma_rrr", 1, 2, sms, ghz, out);
  run<FFMA_IMM>("ffma_imm", 1, 2, sms, ghz, out);
  run<FFMA2>("ffma2 (2 fma per lane-op)", 1, 4, sms, ghz, out);
  run<FADD>("fadd", 1, 1, sms, ghz, out);
  run<FADD2>("fadd2", 1, 2, sms, ghz, out);
  run<PRMT>("prmt", 1, 0, sms, ghz, out);
  run<LOP3>("lop3", 1, 0, sms, ghz, out);
  run<I2FP>("i2fp", 1, 0, sms, ghz, out);
  run<MIX_K1>("2 ffma + 1 prmt (ops counted = 3)", 3, 4.0 / 3, sms, ghz, out);
  run<MIX_K1_F2>("1 ffma2 + 1 prmt (ops counted = 2)", 2, 2, sms, ghz, out);
  return 0;
}
