// channelTcKernel -- the wideband channelizer's RF stage for channels at ARBITRARY frequencies as an int8 GEMM on the
// 5th-generation tensor cores: tcgen05.mma.cta_group::1.kind::i8 with the int32 accumulators in TMEM (sm_100a).
//
// The contraction is channelKernel's (channel_kernels.cuh): for a block of rows (row = one decimation period = D1 samples =
// K = 2 D1 bytes of raw int8 I,Q) and a group of channels,
//     P[row][c][2m + e] = sum_k X[row][k] * B_c[k][2m + e],   B_c[k][.] = h[m D1 + p] exp(j w_c p) / 128  (k = 2p + {I,Q})
// with B in 24-bit fixed point as three signed int8 digits (exact integer MMAs; error <= 2^-24 of the largest entry).  Here
//     A = X          128 rows x K, K-major exactly as the samples lie in memory (row stride K bytes): one TMA box per 128-byte
//                    K-slab, hardware 128-byte swizzle, no repacking
//     B = 240 columns = 5 channels x 3 digits x 16 (8 complex partial sums), K-major, prepared when the channelizer is created
//     D = 128 x 240 int32 in TMEM, two accumulator buffers (2 x 256 of the 512 columns)
// so one tcgen05.mma (M 128, N 240, K 32) does the work of 240 legacy IMMA.16832 fragments, issued by ONE thread.
//
// Warp roles (192 threads, one CTA per SM, persistent over (row tile, channel group) items with the groups fastest, so the
// CTAs running at any moment share a few input tiles and the whole B table in L2):
//   warp 0      TMA producer: per K-slab one box of A (128 rows x 128 B) and one of B (240 rows x 128 B) into a 3-stage ring
//   warp 1      MMA issuer: 4 MMAs per slab, tcgen05.commit frees the stage; after the last slab commits the accumulator
//   warps 2..5  epilogue: tcgen05.ld the accumulator (thread = row), digits -> float, rotate by exp(j w m D1), park the
//               partial sums in shared memory, combine y[k] = sum_m P[k + m][m], demodulate (AM / FM), store coalesced --
//               while the MMA warp already fills the other accumulator buffer.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace b200sdr {

constexpr unsigned kTcRows = 128;           // rows (M) per tile
constexpr unsigned kTcChannels = 5;         // channels per group
constexpr unsigned kTcN = kTcChannels * 48; // 240 accumulator columns: channel x digit x 16
constexpr unsigned kTcStages = 3;
constexpr unsigned kTcABytes = kTcRows * 128u;   // one K-slab of A
constexpr unsigned kTcBBytes = kTcN * 128u;      // one K-slab of B
constexpr unsigned kTcStageBytes = kTcABytes + kTcBBytes;  // 46 KB, a multiple of 1024 (swizzle atoms)
constexpr unsigned kTcThreads = 192;

struct ChannelTcParams {
  float* out;                // demodulated samples, [channel][outStride]
  const float2* rot;         // [channel][8]: exp(j*w_c*m*D1)
  const float* digitScale;   // [channel][3]
  const float* gain;         // [channel]  (FM)
  const int* mod;            // [channel]
  unsigned long long nOut;   // demodulated samples per channel to produce (every row they need is complete in the input)
  unsigned long long outStride;
  unsigned M, kSlabs, numChannels, groups;
  unsigned long long tiles;
  int forceAm;
};

struct ChannelTcSmem {
  unsigned stageOff, parkOff, barOff, total;
};
__host__ __device__ inline ChannelTcSmem channelTcSmemLayout(unsigned M) {
  ChannelTcSmem s;
  s.stageOff = 0;
  s.parkOff = kTcStages * kTcStageBytes;
  s.barOff = s.parkOff + kTcChannels * M * kTcRows * 8u;
  s.total = s.barOff + 256u;
  return s;
}

#ifdef __CUDACC__

__device__ __forceinline__ void tcTmaLoad2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smemAddr(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smemAddr(bar))
               : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart (tools/tc_probe.cu checks these bits)
__device__ __forceinline__ uint64_t tcSmemDesc(const void* tile) {
  uint64_t d = (smemAddr(tile) & 0x3ffffu) >> 4;
  d |= 1ull << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// instruction descriptor, kind::i8: D = s32, A = B = signed 8 bit, both K-major, N / 8 at bit 17, M / 16 at bit 24
__host__ __device__ constexpr uint32_t tcInstrDesc(unsigned m, unsigned n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24); }
__device__ __forceinline__ void tcMma(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmemD), "l"(descA), "l"(descB),
               "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void tcCommit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void tcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcLoad16(uint32_t addr, int (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                 "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(addr));
}

__global__ void __launch_bounds__(kTcThreads, 1)
    channelTcKernel(const ChannelTcParams prm, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ __align__(1024) unsigned char tcsmem[];
  const ChannelTcSmem lay = channelTcSmemLayout(prm.M);
  float2* park = reinterpret_cast<float2*>(tcsmem + lay.parkOff);  // [(c * M + m) * 128 + row]
  uint64_t* full = reinterpret_cast<uint64_t*>(tcsmem + lay.barOff);
  uint64_t* empty = full + kTcStages;
  uint64_t* accFull = empty + kTcStages;
  uint64_t* accEmpty = accFull + 2;
  uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(accEmpty + 2);
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const unsigned M = prm.M, OT = kTcRows - M;
  const unsigned long long items = prm.tiles * prm.groups;

  if (tid == 0) {
    for (unsigned s = 0; s < kTcStages; s++) {
      mbarInit(&full[s], 1);
      mbarInit(&empty[s], 1);
    }
    for (unsigned b = 0; b < 2; b++) {
      mbarInit(&accFull[b], 1);
      mbarInit(&accEmpty[b], 4);
    }
    fenceMbarInit();
  }
  if (warp == 1) {  // the MMA warp owns the tensor memory: all 512 columns (two accumulator buffers of 256)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smemAddr(tmemSlot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcFenceBefore();
  __syncthreads();
  tcFenceAfter();
  const uint32_t tmem = *tmemSlot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      unsigned stage = 0, phase = 0;
      for (unsigned long long item = blockIdx.x; item < items; item += gridDim.x) {
        const unsigned long long tile = item / prm.groups;
        const unsigned group = static_cast<unsigned>(item % prm.groups);
        const int row0 = static_cast<int>(tile * OT);
        for (unsigned slab = 0; slab < prm.kSlabs; slab++) {
          mbarWait(&empty[stage], phase ^ 1u);  // a fresh barrier passes the wait for the phase "before the first"
          unsigned char* sA = tcsmem + lay.stageOff + stage * kTcStageBytes;
          mbarExpectTx(&full[stage], kTcStageBytes);
          tcTmaLoad2d(sA, &mapA, static_cast<int>(slab * 128u), row0, &full[stage]);           // rows past the input: zero-filled
          tcTmaLoad2d(sA + kTcABytes, &mapB, static_cast<int>(slab * 128u), static_cast<int>(group * kTcN), &full[stage]);
          if (++stage == kTcStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = tcInstrDesc(kTcRows, kTcN);
      unsigned stage = 0, phase = 0, buf = 0, accPhase = 0;
      for (unsigned long long item = blockIdx.x; item < items; item += gridDim.x) {
        mbarWait(&accEmpty[buf], accPhase ^ 1u);  // the epilogue has read this accumulator buffer
        tcFenceAfter();
        for (unsigned slab = 0; slab < prm.kSlabs; slab++) {
          mbarWait(&full[stage], phase);
          tcFenceAfter();
          const unsigned char* sA = tcsmem + lay.stageOff + stage * kTcStageBytes;
          const uint64_t dA = tcSmemDesc(sA), dB = tcSmemDesc(sA + kTcABytes);
#pragma unroll
          for (unsigned k4 = 0; k4 < 4; k4++)  // K = 32 per MMA: advance both start addresses by 32 bytes inside the 128-byte swizzle span
            tcMma(tmem + buf * 256u, dA + 2u * k4, dB + 2u * k4, idesc, (slab | k4) != 0u ? 1u : 0u);
          tcCommit(&empty[stage]);  // arrives when the MMAs above have read the stage
          if (++stage == kTcStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tcCommit(&accFull[buf]);
        buf ^= 1u;
        if (buf == 0) accPhase ^= 1u;
      }
    }
  } else {
    // =============================== epilogue: thread <-> row ===============================
    const unsigned row = (warp & 3u) * 32u + lane;  // a warp reads the 32 TMEM lanes of its quarter (warp id mod 4)
    const unsigned et = tid - 64u;                  // 0..127: combine / store index
    unsigned buf = 0, accPhase = 0;
    for (unsigned long long item = blockIdx.x; item < items; item += gridDim.x) {
      const unsigned long long tile = item / prm.groups;
      const unsigned group = static_cast<unsigned>(item % prm.groups);
      const unsigned long long row0 = tile * OT;
      mbarWait(&accFull[buf], accPhase);
      tcFenceAfter();
      const uint32_t base = tmem + ((warp & 3u) * 32u << 16) + buf * 256u;
      for (unsigned c = 0; c < kTcChannels; c++) {
        const unsigned ch = group * kTcChannels + c;
        if (ch >= prm.numChannels) break;  // uniform
        int d0[16], d1[16], d2[16];
        tcLoad16(base + c * 48u, d0);
        tcLoad16(base + c * 48u + 16u, d1);
        tcLoad16(base + c * 48u + 32u, d2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float s0 = prm.digitScale[ch * 3u], s1 = prm.digitScale[ch * 3u + 1u], s2 = prm.digitScale[ch * 3u + 2u];
#pragma unroll
        for (unsigned m = 0; m < 8; m++) {
          if (m < M) {
            const float2 r = prm.rot[ch * 8u + m];
            float2 v;
            v.x = fmaf(static_cast<float>(d2[2 * m]), s2, fmaf(static_cast<float>(d1[2 * m]), s1, static_cast<float>(d0[2 * m]) * s0));
            v.y = fmaf(static_cast<float>(d2[2 * m + 1]), s2, fmaf(static_cast<float>(d1[2 * m + 1]), s1, static_cast<float>(d0[2 * m + 1]) * s0));
            park[(c * M + m) * kTcRows + row] = make_float2(fmaf(-v.y, r.y, v.x * r.x), fmaf(v.y, r.x, v.x * r.y));
          }
        }
      }
      tcFenceBefore();
      __syncwarp();
      if (lane == 0) mbarArrive(&accEmpty[buf]);  // the MMA warp may overwrite this buffer
      asm volatile("bar.sync 1, 128;" ::: "memory");  // every row's partial sums are parked
      const unsigned long long ko = row0 + et;
      if (et < OT && ko < prm.nOut) {
        for (unsigned c = 0; c < kTcChannels; c++) {
          const unsigned ch = group * kTcChannels + c;
          if (ch >= prm.numChannels) break;
          const float2* Pc = park + c * M * kTcRows;
          float2 y = make_float2(0.0f, 0.0f), y1 = make_float2(0.0f, 0.0f);
          for (unsigned m = 0; m < M; m++) {
            const float2 v = Pc[m * kTcRows + et + m], v1 = Pc[m * kTcRows + et + 1u + m];
            y.x += v.x;
            y.y += v.y;
            y1.x += v1.x;
            y1.y += v1.y;
          }
          float o;
          if (prm.mod[ch] == 1 && !prm.forceAm) {  // FM: gain * arg(y[k+1] * conj(y[k]) * exp(j*w*D1))
            const float2 d = make_float2(fmaf(y1.y, y.y, y1.x * y.x), fmaf(y1.y, y.x, -y1.x * y.y));
            const float2 r1 = prm.rot[ch * 8u + 1u];
            o = prm.gain[ch] * atan2f(fmaf(d.y, r1.x, d.x * r1.y), fmaf(-d.y, r1.y, d.x * r1.x));
          } else {
            o = sqrtf(fmaf(y.x, y.x, y.y * y.y));
          }
          prm.out[static_cast<size_t>(ch) * prm.outStride + ko] = o;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the park is free for the next item
      buf ^= 1u;
      if (buf == 0) accPhase ^= 1u;
    }
  }
  tcFenceBefore();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

#endif  // __CUDACC__

}  // namespace b200sdr
