// tc_probe.cu -- smallest end-to-end check of the tcgen05 path this repo's int8 GEMM kernels are built on:
//   TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) -> shared memory -> tcgen05.mma.cta_group::1.kind::i8 (A, B from shared-memory
//   descriptors, K-major, 128-byte swizzle) -> int32 accumulator in TMEM -> tcgen05.ld -> registers -> global,
// compared element by element with the CPU.  C[m][n] = sum_k A[m][k] * B[n][k], M = 128, N = 128, K = 256 (two K-slabs of
// 128 bytes, four MMAs of K = 32 each per slab).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu -lcuda && ./tc_probe
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 128, K = 256, SLAB = 128;  // bytes of K per slab = swizzle span

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni WAIT_DONE;\nbra.uni WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(
          smemAddr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmaLoad2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smemAddr(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smemAddr(bar))
               : "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: start >> 4 | LBO (1, unused for swizzled K-major) << 16 |
// SBO (1024 B between 8-row groups) >> 4 << 32 | version 1 << 46 | layout type 2 (128-byte swizzle) << 61
__device__ __forceinline__ uint64_t umDesc(const void* tile) {
  uint64_t d = (smemAddr(tile) & 0x3ffffu) >> 4;
  d |= 1ull << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// instruction descriptor, kind::i8: D = s32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), both K-major, N >> 3 << 17, M >> 4 << 24
__host__ __device__ constexpr uint32_t umIdesc(int m, int n) { return (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24); }

__global__ void __launch_bounds__(128, 1) tcProbe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                       // 2 slabs x 128 rows x 128 B
  unsigned char* sB = smem + 2 * 128 * SLAB;      // 2 slabs x 128 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * 128 * SLAB);  // [0] loads, [1] mma done
  uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(bars + 2);
  const unsigned tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbarInit(&bars[0], 1);
    mbarInit(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // one warp allocates 128 TMEM columns (one int32 column per n)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smemAddr(tmemSlot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmemSlot;

  if (tid == 0) {
    mbarExpectTx(&bars[0], 4 * 128 * SLAB);
    for (int s = 0; s < 2; s++) {
      tmaLoad2d(sA + s * 128 * SLAB, &mapA, s * SLAB, 0, &bars[0]);
      tmaLoad2d(sB + s * 128 * SLAB, &mapB, s * SLAB, 0, &bars[0]);
    }
    mbarWait(&bars[0], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = umIdesc(M, N);
    for (int s = 0; s < 2; s++)
      for (int k = 0; k < SLAB / 32; k++) {  // K = 32 per MMA: advance the start address by 32 bytes inside the swizzle span
        const uint64_t da = umDesc(sA + s * 128 * SLAB) + ((k * 32) >> 4), db = umDesc(sB + s * 128 * SLAB) + ((k * 32) >> 4);
        const uint32_t accumulate = (s | k) != 0;
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db),
            "r"(idesc), "r"(accumulate)
            : "memory");
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(&bars[1])) : "memory");
  }
  mbarWait(&bars[1], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w owns TMEM lanes 32 w .. 32 w + 31 (= rows m); 32 columns per load
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    const uint32_t addr = tmem + ((warp * 32u) << 16) + c;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; i++) out[tid * N + c + i] = static_cast<int>(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<int8_t> A(M * K), B(N * K);
  srand(1);
  for (auto& v : A) v = static_cast<int8_t>(rand() % 255 - 127);
  for (auto& v : B) v = static_cast<int8_t>(rand() % 255 - 127);
  int8_t *dA, *dB;
  int* dC;
  cudaMalloc(&dA, A.size());
  cudaMalloc(&dB, B.size());
  cudaMalloc(&dC, sizeof(int) * M * N);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  cudaMemset(dC, 0xff, sizeof(int) * M * N);
  auto encode = [&](void* base, int rows, CUtensorMap* map) {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(K)};
    const cuuint32_t box[2] = {SLAB, 128}, es[2] = {1, 1};
    return cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUtensorMap mapA, mapB;
  if (encode(dA, M, &mapA) != CUDA_SUCCESS || encode(dB, N, &mapB) != CUDA_SUCCESS) {
    printf("tensor map encode failed\n");
    return 2;
  }
  const int smem = 4 * 128 * SLAB + 64;
  cudaFuncSetAttribute(tcProbe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  tcProbe<<<1, 128, smem>>>(mapA, mapB, dC);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("kernel failed: %s\n", cudaGetErrorString(e));
    return 3;
  }
  std::vector<int> C(M * N);
  cudaMemcpy(C.data(), dC, sizeof(int) * M * N, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++) {
      int ref = 0;
      for (int k = 0; k < K; k++) ref += static_cast<int>(A[m * K + k]) * static_cast<int>(B[n * K + k]);
      if (ref != C[m * N + n]) {
        if (bad < 5) printf("mismatch at (%d, %d): got %d, want %d\n", m, n, C[m * N + n], ref);
        bad++;
      }
    }
  printf("tc_probe: %ld mismatches of %d\n", bad, M * N);
  return bad == 0 ? 0 : 1;
}
