// channelKernel -- the wideband channelizer's RF stage (BASELINE config C5): ONE pass over the int8 IQ input produces
// the demodulated (AM / FM) stream of a GROUP of channels.  For a block of rows (one row = D1 consecutive samples = one
// decimation period) and NC channels the convert + mix + polyphase-FIR work is the int8 GEMM
//
//     P[row][c][2m + e] = sum_k X[row][k] * B_c[k][2m + e],      k = 2p + (0: I, 1: Q),  e = (0: re, 1: im),
//
// with X the raw int8 input exactly as it lies in memory (row stride 2*D1 bytes, K = 2*D1: 1280 for C5) and B_c the
// channel's mixer-rotated taps g_c[p][m] = h[m*D1 + p] * exp(j*w_c*p) / 128 in 24-bit fixed point (three signed int8
// digits, so three exact IMMA.16832.S8.S8 per k-step; see chain_kernels.cuh for the derivation and the error bound).
// The per-row carrier exp(j*w_c*q*D1) is never applied: |.| and the FM discriminator do not depend on it.
//
//   grid  = (channel groups, row tiles)   -- channel groups vary fastest, so the CTAs that share an input tile run
//                                            back to back and the tile is served from L2 after its first read
//   CTA   = W warps x 32 rows; K is streamed in 32-byte steps through a 3-stage cp.async ring (A: rows x 32 B of raw
//           samples, B: the group's fragment-ordered digits of that k-step, contiguous in HBM)
//   warp  = 2 m-tiles x (NC * NTC) n-tiles x 3 digits accumulators in registers
//   epilogue: digits -> float, rotate, park P in shared memory (re-using the ring), then one thread per output row
//           combines y[k] = sum_m P[k+m][m], demodulates and stores to out[channel][k] (coalesced).
#pragma once

#include "common.cuh"

namespace b200sdr {

struct ChannelParams {
  const unsigned char* in;   // int8 I,Q pairs
  float* out;                // demodulated samples, [channel][outStride]
  const unsigned* bFrag;     // [group][k-step][n-tile (NC*NTC)][digit 3][half 2][lane 32] packed int8x4
  const float2* rot;         // [channel][8]: exp(j*w_c*m*D1)
  const float* digitScale;   // [channel][3]
  const float* gain;         // [channel]  (FM)
  const int* mod;            // [channel]  kModAm / kModFm
  unsigned long long nIn;    // valid input samples
  unsigned long long nOut;   // demodulated samples per channel
  unsigned long long outStride;
  unsigned D1, M, kSteps;
  unsigned numChannels;
  int forceAm;               // demodulate every channel as AM whatever `mod` says (the AM tail pass of a mixed AM/FM set)
};

constexpr int kChanNC = 4;        // channels per CTA
constexpr int kChanStages = 3;    // cp.async ring depth
constexpr unsigned kChanARow = 48;  // bytes per row of an A chunk in shared memory (32 used; 48 keeps fragment reads conflict-free)

#ifdef __CUDACC__

__device__ __forceinline__ void cpAsync16(void* smemDst, const void* gmemSrc, bool valid) {
  const unsigned bytes = valid ? 16u : 0u;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smemAddr(smemDst)), "l"(gmemSrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpAsyncWait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void imma16832c(int (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// NTC = n-tiles (of 8 columns) per channel: 1 for M <= 4 partial sums, 2 for M <= 8
template <int NTC>
__global__ void __launch_bounds__(256) channelKernel(const ChannelParams prm) {
  constexpr unsigned NC = kChanNC, NTILES = NC * NTC;
  constexpr unsigned B_CHUNK_WORDS = NTILES * 3u * 64u;  // words of one k-step's fragments for the group
  extern __shared__ __align__(128) unsigned char smem[];

  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const unsigned W = blockDim.x >> 5, rowsTile = W * 32u;
  const unsigned g = lane >> 2, t = lane & 3u;
  const unsigned D = prm.D1, M = prm.M, KS = prm.kSteps;
  const unsigned OT = rowsTile - M;  // output rows per tile: M-1 rows of FIR halo + 1 row for the FM discriminator
  const unsigned group = blockIdx.x;
  const unsigned long long row0 = static_cast<unsigned long long>(blockIdx.y) * OT;
  const unsigned rowBytes = 2u * D;
  const unsigned long long totalBytes = prm.nIn * 2ull;

  const unsigned aChunkBytes = rowsTile * kChanARow;
  const unsigned stageBytes = aChunkBytes + B_CHUNK_WORDS * 4u;
  const unsigned* bGroup = prm.bFrag + static_cast<size_t>(group) * KS * B_CHUNK_WORDS;

  auto issue = [&](unsigned ks, unsigned stage) {
    unsigned char* sa = smem + stage * stageBytes;
    unsigned* sb = reinterpret_cast<unsigned*>(sa + aChunkBytes);
    // A: two 16-byte pieces per row
    for (unsigned i = tid; i < rowsTile * 2u; i += blockDim.x) {
      const unsigned r = i >> 1, piece = i & 1u;
      const unsigned long long off = (row0 + r) * rowBytes + ks * 32u + piece * 16u;
      const bool valid = ks * 32u + piece * 16u < rowBytes && off + 16u <= totalBytes;
      cpAsync16(sa + r * kChanARow + piece * 16u, prm.in + (valid ? off : 0ull), valid);
    }
    // B: contiguous chunk
    const unsigned* src = bGroup + static_cast<size_t>(ks) * B_CHUNK_WORDS;
    for (unsigned i = tid; i < B_CHUNK_WORDS / 4u; i += blockDim.x) cpAsync16(sb + i * 4u, src + i * 4u, true);
  };

  int acc[2][NTILES][3][4];
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (unsigned n = 0; n < NTILES; n++)
#pragma unroll
      for (int d = 0; d < 3; d++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[j][n][d][e] = 0;

  // ---- k loop: 3-stage ring ---------------------------------------------------------------------------------
#pragma unroll
  for (unsigned s = 0; s < kChanStages - 1; s++) {
    if (s < KS) issue(s, s);
    cpAsyncCommit();
  }
#pragma unroll 1
  for (unsigned ks = 0; ks < KS; ks++) {
    cpAsyncWait<kChanStages - 2>();
    __syncthreads();  // chunk ks has landed for every thread; everyone is done with the chunk the next issue overwrites
    if (ks + kChanStages - 1 < KS) issue(ks + kChanStages - 1, (ks + kChanStages - 1) % kChanStages);
    cpAsyncCommit();

    const unsigned char* sa = smem + (ks % kChanStages) * stageBytes;
    const unsigned* sb = reinterpret_cast<const unsigned*>(sa + aChunkBytes);
    unsigned a[2][4];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const unsigned char* lo = sa + (warp * 32u + j * 16u + g) * kChanARow + t * 4u;
      const unsigned char* hi = lo + 8u * kChanARow;
      a[j][0] = *reinterpret_cast<const unsigned*>(lo);
      a[j][1] = *reinterpret_cast<const unsigned*>(hi);
      a[j][2] = *reinterpret_cast<const unsigned*>(lo + 16u);
      a[j][3] = *reinterpret_cast<const unsigned*>(hi + 16u);
    }
#pragma unroll
    for (unsigned n = 0; n < NTILES; n++) {
#pragma unroll
      for (unsigned d = 0; d < 3; d++) {
        const unsigned b0 = sb[((n * 3u + d) * 2u) * 32u + lane], b1 = sb[((n * 3u + d) * 2u + 1u) * 32u + lane];
#pragma unroll
        for (int j = 0; j < 2; j++) imma16832c(acc[j][n][d], a[j][0], a[j][1], a[j][2], a[j][3], b0, b1);
      }
    }
  }
  cpAsyncWait<0>();
  __syncthreads();  // the ring is free: re-use it for the partial sums

  // ---- digits -> float, rotate, park: P[c][m][row] ----------------------------------------------------------
  float2* P = reinterpret_cast<float2*>(smem);  // [(c*M + m) * rowsTile + row]
#pragma unroll
  for (unsigned n = 0; n < NTILES; n++) {
    const unsigned c = n / NTC, m = (n % NTC) * 4u + t;
    const unsigned ch = group * NC + c;
    if (m < M && ch < prm.numChannels) {
      const float s0 = prm.digitScale[ch * 3u], s1 = prm.digitScale[ch * 3u + 1u], s2 = prm.digitScale[ch * 3u + 2u];
      const float2 r = prm.rot[ch * 8u + m];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        float2 lo, hi;
        lo.x = fmaf(static_cast<float>(acc[j][n][2][0]), s2, fmaf(static_cast<float>(acc[j][n][1][0]), s1, static_cast<float>(acc[j][n][0][0]) * s0));
        lo.y = fmaf(static_cast<float>(acc[j][n][2][1]), s2, fmaf(static_cast<float>(acc[j][n][1][1]), s1, static_cast<float>(acc[j][n][0][1]) * s0));
        hi.x = fmaf(static_cast<float>(acc[j][n][2][2]), s2, fmaf(static_cast<float>(acc[j][n][1][2]), s1, static_cast<float>(acc[j][n][0][2]) * s0));
        hi.y = fmaf(static_cast<float>(acc[j][n][2][3]), s2, fmaf(static_cast<float>(acc[j][n][1][3]), s1, static_cast<float>(acc[j][n][0][3]) * s0));
        lo = make_float2(fmaf(-lo.y, r.y, lo.x * r.x), fmaf(lo.y, r.x, lo.x * r.y));
        hi = make_float2(fmaf(-hi.y, r.y, hi.x * r.x), fmaf(hi.y, r.x, hi.x * r.y));
        const unsigned row = warp * 32u + j * 16u + g;
        P[(c * M + m) * rowsTile + row] = lo;
        P[(c * M + m) * rowsTile + row + 8u] = hi;
      }
    }
  }
  __syncthreads();

  // ---- one thread per output row: combine, demodulate, store -------------------------------------------------
  for (unsigned k = tid; k < OT; k += blockDim.x) {
    const unsigned long long ko = row0 + k;
    if (ko >= prm.nOut) break;
#pragma unroll
    for (unsigned c = 0; c < NC; c++) {
      const unsigned ch = group * NC + c;
      if (ch >= prm.numChannels) break;
      const float2* Pc = P + c * M * rowsTile;
      float2 y = make_float2(0.0f, 0.0f), y1 = make_float2(0.0f, 0.0f);
      for (unsigned m = 0; m < M; m++) {
        const float2 v = Pc[m * rowsTile + k + m], v1 = Pc[m * rowsTile + k + 1u + m];
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
}

#endif  // __CUDACC__

}  // namespace b200sdr
