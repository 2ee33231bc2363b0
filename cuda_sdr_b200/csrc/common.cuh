// Shared device/host helpers for the sm_100a kernels of the DSP hot path.
#pragma once

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace b200sdr {

constexpr int kSmCount = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

extern std::atomic<uint64_t> g_launchCount;  // kernels launched by this library (bench.py gpu_launches)

// RAII: make `device` current for the duration of one API call, like the reference's
// CudaDevicePushPop (include/gpusdrpipeline/util/CudaDevicePushPop.h:27-54).
struct DeviceGuard {
  int previous = -1;
  cudaError_t status = cudaSuccess;
  explicit DeviceGuard(int device) {
    status = cudaGetDevice(&previous);
    if (status == cudaSuccess && previous != device) status = cudaSetDevice(device);
  }
  ~DeviceGuard() {
    int now = -1;
    if (previous >= 0 && cudaGetDevice(&now) == cudaSuccess && now != previous) cudaSetDevice(previous);
  }
};

inline cudaError_t launchStatus() {
  g_launchCount.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

// Grid for a grid-stride element-wise kernel: enough CTAs to cover n, capped at 8 resident CTAs/SM.
inline unsigned elementwiseGrid(size_t workItems, unsigned threads) {
  size_t blocks = (workItems + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(kSmCount) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return static_cast<unsigned>(blocks);
}

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smemAddr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 128-bit streaming loads/stores: data is touched once, keep it out of L1.
__device__ __forceinline__ uint4 ldStream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldStream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stStream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// Four int8 packed in one 32-bit word -> four exact floats (NOT yet scaled by 1/128).
// Bias each byte to unsigned (x+128), drop it into the mantissa of 2^23 with one PRMT, subtract
// 2^23+128 with one FADD: exact, and uses the ALU and FMA pipes instead of the quarter-rate XU I2F.
__device__ __forceinline__ void int8x4ToFloat(uint32_t w, float& a, float& b, float& c, float& d) {
  constexpr uint32_t kMagic = 0x4B000000u;  // 2^23
  constexpr float kBias = 8388736.0f;       // 2^23 + 128
  w ^= 0x80808080u;
  a = __uint_as_float(__byte_perm(w, kMagic, 0x7650)) - kBias;
  b = __uint_as_float(__byte_perm(w, kMagic, 0x7651)) - kBias;
  c = __uint_as_float(__byte_perm(w, kMagic, 0x7652)) - kBias;
  d = __uint_as_float(__byte_perm(w, kMagic, 0x7653)) - kBias;
}

// mbarrier + 1-D TMA bulk copy (global -> shared), sm_90+/sm_100a.
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void fenceMbarInit() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmaBulkLoad(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(smemDst)),
      "l"(gmemSrc), "r"(bytes), "r"(smemAddr(bar))
      : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smemAddr(bar)),
      "r"(parity)
      : "memory");
}

#endif  // __CUDACC__

}  // namespace b200sdr
