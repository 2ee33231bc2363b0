// chainKernel -- the whole int8/cf32 -> [mix] -> decimating FIR -> AM/FM demod -> audio FIR chain in ONE persistent,
// warp-specialised kernel (sm_100a).  HBM traffic is the algorithmic minimum: every input byte is read once (plus a
// small halo per CTA), only the final audio samples are written; the demodulated stream never leaves shared memory.
// See the comment above the kernel for the division of labour between the warps.
#pragma once

#include "fir_kernels.cuh"

// cycle counters inside the kernel cost issue slots; they are compiled in only for tools (-DB200SDR_CHAIN_PROFILE_BUILD)
#ifdef B200SDR_CHAIN_PROFILE_BUILD
#define B200SDR_PROF(prm__) ((prm__).prof != nullptr)
#else
#define B200SDR_PROF(prm__) false
#endif

namespace b200sdr {

struct ChainParams {
  const void* in;          // int8 pairs or float2
  float* out;              // audio outputs
  const float* tapTable;   // hT[D1][TS], pre-scaled
  const float2* mixTable;  // W[D1] (scaled by inScale); unused when !MIX
  const float2* rotTable;  // exp(j*w*m*D1), m <= TS; unused when !MIX
  const float* taps2;      // T2 audio taps
  unsigned long long nIn;     // valid input elements
  unsigned long long nAudio;  // audio outputs to produce
  unsigned T1, D1, M, T2, D2;
  unsigned stages;         // TMA ring depth (<= 8)
  unsigned audioWarps;     // warps of the CTA that run the audio FIR (blockDim = 32 * (computeWarps + audioWarps))
  unsigned dmCapacity;     // floats per demod line (>= outputs per tile + T2)
  int mod;                 // kModAm / kModFm
  float gain;
  // tensor-core route (int8 input): B fragments of the fixed-point tap/mixer matrix and the digit weights
  const unsigned* bFrag;   // [digit 3][n-tile][k-step][half 2][lane 32] packed int8x4
  unsigned kSteps;         // ceil(2*D1 / 32)
  float digitScale[3];     // value = sum_d acc_d * digitScale[d]
  unsigned long long* prof;  // optional (tools only): per-warp cycle counters [grid][warps][6] = wait-tile, main loop, rest, wait-line, prologue, total
};

#ifdef __CUDACC__

__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// RPT rows per thread, MP partial sums per row: thread rows are firstRow + i*rowStride of the row block at `tile`
// (raw samples staged in shared memory).  Written so that the work of the RPT rows is visibly independent: the next
// 16-byte chunk of every row is fetched before the current one is consumed, and each stage (convert, mix, accumulate)
// runs over all rows before the next stage starts, which lets ptxas interleave the dependent chains of different rows.
template <int ELEM, bool MIX, int MP, int RPT, int CONV>
__device__ __forceinline__ void tilePartialSums(const unsigned char* tile, const float* hT, const float4* W, unsigned D, unsigned firstRow,
                                                unsigned rowStride, float2 (&acc)[RPT][MP]) {
  constexpr int ES = ElemTraits<ELEM>::kBytes;
  constexpr int VEC = ElemTraits<ELEM>::kVec;
#pragma unroll
  for (int i = 0; i < RPT; i++)
#pragma unroll
    for (int m = 0; m < MP; m++) acc[i][m] = make_float2(0.0f, 0.0f);
  const unsigned rowBytes = D * ES;
  const uint4* row[RPT];
#pragma unroll
  for (int i = 0; i < RPT; i++) row[i] = reinterpret_cast<const uint4*>(tile + (firstRow + i * rowStride) * rowBytes);
  const unsigned chunks = D / VEC;

  uint4 v[RPT];
#pragma unroll
  for (int i = 0; i < RPT; i++) v[i] = row[i][0];

#pragma unroll 1
  for (unsigned c = 0; c < chunks; c++) {
    uint4 nxt[RPT];
    const unsigned cn = c + 1 < chunks ? c + 1 : c;  // the last iteration re-reads its own chunk (harmless, branch-free)
#pragma unroll
    for (int i = 0; i < RPT; i++) nxt[i] = row[i][cn];
    const unsigned p = c * VEC;

    if constexpr (ELEM == kElemInt8Complex) {
#pragma unroll
      for (int s = 0; s < 4; s++) {  // word s holds samples p+2s and p+2s+1 as I,Q,I,Q bytes
        float h0[MP], h1[MP];
        loadTapRow<MP>(hT, p + 2 * s, h0);
        loadTapRow<MP>(hT, p + 2 * s + 1, h1);
        float4 w0, w1;
        if constexpr (MIX) {
          w0 = W[p + 2 * s];
          w1 = W[p + 2 * s + 1];
        }
        float2 z0[RPT], z1[RPT];
#pragma unroll
        for (int i = 0; i < RPT; i++) {
          const uint32_t word = s == 0 ? v[i].x : s == 1 ? v[i].y : s == 2 ? v[i].z : v[i].w;
          convertWord<CONV>(word, z0[i], z1[i]);
        }
        if constexpr (MIX) {
#pragma unroll
          for (int i = 0; i < RPT; i++) {
            z0[i] = cmulPacked(z0[i], w0);
            z1[i] = cmulPacked(z1[i], w1);
          }
        }
#pragma unroll
        for (int m = 0; m < MP; m++) {
#pragma unroll
          for (int i = 0; i < RPT; i++) acc[i][m] = axpy2(h0[m], z0[i], acc[i][m]);
#pragma unroll
          for (int i = 0; i < RPT; i++) acc[i][m] = axpy2(h1[m], z1[i], acc[i][m]);
        }
      }
    } else {
      float h0[MP], h1[MP];
      loadTapRow<MP>(hT, p, h0);
      loadTapRow<MP>(hT, p + 1, h1);
      float4 w0, w1;
      if constexpr (MIX) {
        w0 = W[p];
        w1 = W[p + 1];
      }
      float2 z0[RPT], z1[RPT];
#pragma unroll
      for (int i = 0; i < RPT; i++) {
        z0[i] = make_float2(__uint_as_float(v[i].x), __uint_as_float(v[i].y));
        z1[i] = make_float2(__uint_as_float(v[i].z), __uint_as_float(v[i].w));
        if constexpr (MIX) {
          z0[i] = cmulPacked(z0[i], w0);
          z1[i] = cmulPacked(z1[i], w1);
        }
      }
#pragma unroll
      for (int m = 0; m < MP; m++) {
#pragma unroll
        for (int i = 0; i < RPT; i++) acc[i][m] = axpy2(h0[m], z0[i], acc[i][m]);
#pragma unroll
        for (int i = 0; i < RPT; i++) acc[i][m] = axpy2(h1[m], z1[i], acc[i][m]);
      }
    }
#pragma unroll
    for (int i = 0; i < RPT; i++) v[i] = nxt[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core route for int8 input.  Once the per-row carrier is hoisted (AM/FM outputs do not depend on it) the
// convert + mix + polyphase FIR of a row block IS a dense contraction
//     P[row][2m + c] = sum_k X[row][k] * B[k][2m + c],   k = 2p + (0: I, 1: Q),  c = (0: re, 1: im),
// with X the raw int8 tile exactly as it lies in shared memory (row stride 2*D1 bytes) and B the mixer-rotated taps
// g[p][m] = h[m*D1 + p] * exp(j*w*p) / 128 laid out as B[2p][2m] = re g, B[2p+1][2m] = -im g, B[2p][2m+1] = im g,
// B[2p+1][2m+1] = re g.  B is held in 24-bit fixed point as three signed int8 digits, so three exact int8 x int8 -> int32
// MMAs (IMMA.16832.S8.S8) per k-step replace ~420 dispatch cycles of CUDA-core convert/mix/FMA work per row; the
// int8 samples are never converted.  Fixed-point error is <= 2^-24 of the largest matrix entry per tap.
//
// Fragment layout of mma.sync.m16n8k32 (g = lane / 4, t = lane % 4): A regs = rows {g, g+8} x k {4t.., 16+4t..};
// B regs = k {4t.., 16+4t..} x column g; C regs = rows {g, g+8} x columns {2t, 2t+1} = one complex partial sum m = t.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void imma16832(int (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Writes the (rotated) partial sums of the warp's RB = 32*RPT rows to part[m*RB + row].  G m-tiles of 16 rows are
// in flight at once (G*NT*3 independent accumulator chains), so the k-loop runs at IMMA throughput, not IMMA latency.
template <int MP, int RPT, bool MIX>
__device__ __forceinline__ void tileMmaPartials(const unsigned char* block, const unsigned* bFrag, unsigned kSteps, unsigned D, unsigned M,
                                                const float2* rot, float s0, float s1, float s2, float2* part, unsigned lane) {
  constexpr unsigned RB = 32u * RPT;
  constexpr unsigned NT = (2u * MP + 7u) / 8u;
  constexpr unsigned MT = RB / 16u;
  constexpr unsigned G = NT == 1 ? (MT < 8u ? MT : 8u) : (MT < 4u ? MT : 4u);
  const unsigned g = lane >> 2, t = lane & 3u;
  const unsigned rowBytes = 2u * D;
#pragma unroll 1
  for (unsigned mt0 = 0; mt0 < MT; mt0 += G) {
    int acc[G][NT][3][4];
#pragma unroll
    for (unsigned j = 0; j < G; j++)
#pragma unroll
      for (unsigned n = 0; n < NT; n++)
#pragma unroll
        for (int d = 0; d < 3; d++)
#pragma unroll
          for (int e = 0; e < 4; e++) acc[j][n][d][e] = 0;
    const unsigned char* r0 = block + (mt0 * 16u + g) * rowBytes + t * 4u;
#pragma unroll 1
    for (unsigned ks = 0; ks < kSteps; ks++) {
      unsigned b[NT][3][2];
#pragma unroll
      for (unsigned n = 0; n < NT; n++)
#pragma unroll
        for (unsigned d = 0; d < 3; d++) {
          const unsigned* src = bFrag + (((d * NT + n) * kSteps + ks) * 2u) * 32u + lane;
          b[n][d][0] = src[0];
          b[n][d][1] = src[32];
        }
      unsigned a[G][4];
#pragma unroll
      for (unsigned j = 0; j < G; j++) {
        const unsigned char* lo = r0 + j * 16u * rowBytes + ks * 32u;
        const unsigned char* hi = lo + 8u * rowBytes;
        a[j][0] = *reinterpret_cast<const unsigned*>(lo);
        a[j][1] = *reinterpret_cast<const unsigned*>(hi);
        a[j][2] = *reinterpret_cast<const unsigned*>(lo + 16u);
        a[j][3] = *reinterpret_cast<const unsigned*>(hi + 16u);
      }
#pragma unroll
      for (unsigned j = 0; j < G; j++)
#pragma unroll
        for (unsigned n = 0; n < NT; n++)
#pragma unroll
          for (unsigned d = 0; d < 3; d++) imma16832(acc[j][n][d], a[j][0], a[j][1], a[j][2], a[j][3], b[n][d][0], b[n][d][1]);
    }
    // digits -> float, rotate by exp(j*w*m*D1), park in the warp's scratch
#pragma unroll
    for (unsigned n = 0; n < NT; n++) {
      const unsigned m = n * 4u + t;
      if (m < M) {
        float2 r = make_float2(1.0f, 0.0f);
        if constexpr (MIX) r = rot[m];  // rot[0] = 1
#pragma unroll
        for (unsigned j = 0; j < G; j++) {
          float2 lo, hi;  // rows g and g+8 of m-tile mt0+j
          lo.x = fmaf(static_cast<float>(acc[j][n][2][0]), s2, fmaf(static_cast<float>(acc[j][n][1][0]), s1, static_cast<float>(acc[j][n][0][0]) * s0));
          lo.y = fmaf(static_cast<float>(acc[j][n][2][1]), s2, fmaf(static_cast<float>(acc[j][n][1][1]), s1, static_cast<float>(acc[j][n][0][1]) * s0));
          hi.x = fmaf(static_cast<float>(acc[j][n][2][2]), s2, fmaf(static_cast<float>(acc[j][n][1][2]), s1, static_cast<float>(acc[j][n][0][2]) * s0));
          hi.y = fmaf(static_cast<float>(acc[j][n][2][3]), s2, fmaf(static_cast<float>(acc[j][n][1][3]), s1, static_cast<float>(acc[j][n][0][3]) * s0));
          if constexpr (MIX) {
            lo = cmulf(lo, r);
            hi = cmulf(hi, r);
          }
          part[m * RB + (mt0 + j) * 16u + g] = lo;
          part[m * RB + (mt0 + j) * 16u + g + 8u] = hi;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-specialised persistent kernel.  blockDim = 32 * (NW + 1): NW compute warps + 1 audio warp.
//
//   compute warp w : owns the row block [w*OTW, w*OTW + 32*RPT) of every tile (blocks overlap by M-1+fm rows, so a
//                    warp needs no row of another warp): main loop -> partial sums exchanged through the warp's own
//                    shared-memory scratch (__syncwarp only) -> demodulate -> OTW samples into the demod line.
//                    The LAST warp to finish a tile slot (a shared-memory counter tells which) refills it by TMA at
//                    once -- no producer warp, nobody waits to issue a copy.
//   audio warp     : waits for a demod line to be complete (mbarrier, NW arrivals), computes every audio output whose
//                    window is complete (one lane per output, 128-bit tap reads), carries the tail into the other
//                    line and hands the line back (mbarrier).
// There is no CTA-wide barrier in the steady state: tiles, lines and slots are handed over by mbarriers.
// ------------------------------------------------------------------------------------------------
struct ChainSmem2 {
  unsigned mixOff, rotOff, tapOff, taps2Off, exchOff, exchBytesPerWarp, bFragOff, dmOff, tileOff, tileBytes, slotStride, total;
};

// barriers: full[8] at 0, dmFull[2] at 64, dmEmpty[2] at 80, slot counters done[8] (u32) at 96
__host__ __device__ inline ChainSmem2 chainSmemLayout2(unsigned D, unsigned TS, unsigned M, unsigned rpt, unsigned computeWarps, unsigned elemBytes,
                                                       bool fm, unsigned T2, unsigned dmCapacity, unsigned stages, bool mma = false,
                                                       unsigned bFragWords = 0) {
  ChainSmem2 s;
  unsigned off = 128;
  s.mixOff = off;
  off += D * 16;
  s.rotOff = off;
  off += (TS + 1) * 8;
  off = (off + 15u) & ~15u;
  s.tapOff = off;
  off += D * TS * 4;
  off = (off + 15u) & ~15u;
  s.taps2Off = off;
  off += ((T2 + 3u) & ~3u) * 4;
  s.exchOff = off;
  // CUDA-core route: partial sums m >= 1 (+ the FM line); tensor-core route: all M partial sums (+ the FM line)
  s.exchBytesPerWarp = ((mma ? M : (M > 0 ? M - 1 : 0)) + (fm ? 1u : 0u)) * 32u * rpt * 8u;
  off += s.exchBytesPerWarp * computeWarps;
  s.bFragOff = off;
  off += mma ? bFragWords * 4u : 0u;
  off = (off + 15u) & ~15u;
  s.dmOff = off;
  off += 2u * dmCapacity * 4;
  off = (off + 127u) & ~127u;
  s.tileOff = off;
  const unsigned outPerWarp = 32u * rpt - (M - 1) - (fm ? 1u : 0u);
  const unsigned tileRows = computeWarps * outPerWarp + (M - 1) + (fm ? 1u : 0u);
  s.tileBytes = tileRows * D * elemBytes;
  s.slotStride = (s.tileBytes + 32u + 127u) & ~127u;  // 32 B of slack: the last k-step of the last row may read past the row
  off += stages * s.slotStride;
  s.total = off;
  return s;
}

template <int ELEM, bool MIX, int MP, int RPT, int CONV>
__global__ void __launch_bounds__(RPT >= 4 ? 192 : RPT == 2 ? 320 : 576, RPT >= 2 ? 2 : 1) chainKernel(const ChainParams prm) {
  constexpr int ES = ElemTraits<ELEM>::kBytes;
  constexpr int TS = tapStride(MP);
  constexpr unsigned RB = 32u * RPT;  // rows per warp block
  extern __shared__ __align__(128) unsigned char smem[];

  const unsigned NA = prm.audioWarps;
  const unsigned NW = blockDim.x / 32u - NA;
  const unsigned D = prm.D1, M = prm.M, S = prm.stages;
  const unsigned T2 = prm.T2, D2 = prm.D2;
  const bool fm = prm.mod == kModFm;
  const unsigned OTW = RB - (M - 1) - (fm ? 1u : 0u);  // demod outputs per warp block
  const unsigned OT = NW * OTW;                         // demod outputs per tile
  constexpr bool MMA = CONV == 2;
  constexpr unsigned NT = (2u * MP + 7u) / 8u;  // n-tiles of 8 columns: 2 real columns per partial sum
  const ChainSmem2 lay = chainSmemLayout2(D, TS, M, RPT, NW, ES, fm, T2, prm.dmCapacity, S, MMA, 3u * NT * prm.kSteps * 64u);
  const unsigned* bFrag = reinterpret_cast<const unsigned*>(smem + lay.bFragOff);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* dmFull = reinterpret_cast<uint64_t*>(smem + 64);
  uint64_t* dmEmpty = reinterpret_cast<uint64_t*>(smem + 80);
  unsigned* slotDone = reinterpret_cast<unsigned*>(smem + 96);
  float4* W = reinterpret_cast<float4*>(smem + lay.mixOff);
  float2* rot = reinterpret_cast<float2*>(smem + lay.rotOff);
  float* hT = reinterpret_cast<float*>(smem + lay.tapOff);
  float* h2 = reinterpret_cast<float*>(smem + lay.taps2Off);
  float* dm = reinterpret_cast<float*>(smem + lay.dmOff);
  unsigned char* tiles = smem + lay.tileOff;
  const unsigned tileBytes = lay.tileBytes, slotStride = lay.slotStride;

  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const long long tEntry = B200SDR_PROF(prm) ? clock64() : 0;

  // ---- this CTA's run of audio outputs ---------------------------------------------------------------
  const unsigned long long per = prm.nAudio / gridDim.x, extra = prm.nAudio % gridDim.x;
  const unsigned long long a0 = blockIdx.x * per + (blockIdx.x < extra ? blockIdx.x : extra);
  const unsigned long long cnt = per + (blockIdx.x < extra ? 1 : 0);
  if (cnt == 0) return;
  const unsigned long long row0 = a0 * D2;
  const unsigned long long needDemod = (cnt - 1) * D2 + T2;
  const unsigned nTiles = static_cast<unsigned>((needDemod + OT - 1) / OT);
  const unsigned long long totalBytes = prm.nIn * ES;
  const unsigned char* gin = static_cast<const unsigned char*>(prm.in);

  auto tileStart = [&](unsigned t) { return (row0 + static_cast<unsigned long long>(t) * OT) * D * ES; };
  auto tileAvail = [&](unsigned t) -> unsigned {
    const unsigned long long start = tileStart(t);
    if (start >= totalBytes) return 0u;
    const unsigned long long left = totalBytes - start;
    return left < tileBytes ? static_cast<unsigned>(left) : tileBytes;
  };
  // one thread: arm the slot's barrier and start the bulk copy of tile t
  auto issueTileInto = [&](unsigned t, unsigned slot) {
    const unsigned bulk = tileAvail(t) & ~15u;
    mbarExpectTx(&full[slot], bulk);
    if (bulk) tmaBulkLoad(tiles + slot * slotStride, gin + tileStart(t), bulk, &full[slot]);
  };
  auto issueTile = [&](unsigned t) { issueTileInto(t, t % S); };
  // the bytes of tile t the bulk copy does not cover (a <16 B remainder, and zeros past the end of the input); `who`
  // of `count` cooperating threads
  auto fillTileTail = [&](unsigned t, unsigned who, unsigned count) {
    const unsigned avail = tileAvail(t);
    if (avail == tileBytes) return;  // every tile but the last one or two of the whole input
    unsigned char* dst = tiles + (t % S) * slotStride;
    const unsigned long long start = tileStart(t);
    for (unsigned b = (avail & ~15u) + who; b < tileBytes; b += count) dst[b] = b < avail ? gin[start + b] : 0;
  };

  // ---- prologue (the only CTA-wide barrier) --------------------------------------------------------------
  if (tid == 0) {
    for (unsigned s = 0; s < S; s++) {
      mbarInit(&full[s], 1);
      slotDone[s] = 0;
    }
    for (unsigned i = 0; i < 2; i++) {
      mbarInit(&dmFull[i], NW);
      mbarInit(&dmEmpty[i], NA);
    }
    fenceMbarInit();
  }
  for (unsigned t = 0; t < S && t < nTiles; t++) fillTileTail(t, tid, blockDim.x);
  if constexpr (MMA) {
    unsigned* dst = reinterpret_cast<unsigned*>(smem + lay.bFragOff);
    for (unsigned i = tid; i < 3u * NT * prm.kSteps * 64u; i += blockDim.x) dst[i] = prm.bFrag[i];
  }
  for (unsigned i = tid; i < D * TS; i += blockDim.x) hT[i] = prm.tapTable[i];
  for (unsigned i = tid; i < ((T2 + 3u) & ~3u); i += blockDim.x) h2[i] = i < T2 ? prm.taps2[i] : 0.0f;
  if constexpr (MIX) {
    for (unsigned p = tid; p < D; p += blockDim.x) {
      const float2 w = prm.mixTable[p];
      W[p] = make_float4(w.x, w.y, -w.y, w.x);
    }
    if (tid <= TS) rot[tid] = prm.rotTable[tid];
  }
  __syncthreads();
  if (tid == 0) {
    fenceProxyAsync();
    for (unsigned t = 0; t < S && t < nTiles; t++) issueTile(t);
  }

  // carry / done bookkeeping is a pure function of the tile index: every warp tracks it on its own
  unsigned carry = 0;
  unsigned long long done = 0;
  const unsigned otDiv = OT / D2, otRem = OT % D2;
  // audio outputs that become computable once a tile's OT samples are appended to `carry` leftover samples
  auto outputsReady = [&](unsigned carryNow) -> unsigned {
    const unsigned len = carryNow + OT;
    unsigned nA;
    if (carryNow + D2 >= T2 && carryNow < T2) {
      nA = otDiv + (carryNow + otRem >= T2 ? 1u : 0u);  // steady state: no division
    } else {
      nA = len >= T2 ? (len - T2) / D2 + 1 : 0;
    }
    const unsigned long long left = cnt - done;
    return static_cast<unsigned long long>(nA) > left ? static_cast<unsigned>(left) : nA;
  };

  if (warp < NW) {
    // =========================== compute warps ===========================
    float2* part = reinterpret_cast<float2*>(smem + lay.exchOff + warp * lay.exchBytesPerWarp);  // [(m-1)*RB + row]
    float2* sums = part + (MMA ? M : (M > 0 ? (M - 1) : 0)) * RB;                                  // FM only
    float2 rot1 = make_float2(1.0f, 0.0f);
    if constexpr (MIX) rot1 = rot[1];

    long long cWait = 0, cMain = 0, cRest = 0, cLine = 0;
    const long long tLoop = B200SDR_PROF(prm) ? clock64() : 0;
    unsigned slot = 0, slotPhase = 0;  // t % S and (t / S) & 1 without the divisions
    for (unsigned t = 0; t < nTiles; t++) {
      const long long t0 = B200SDR_PROF(prm) ? clock64() : 0;
      mbarWait(&full[slot], slotPhase);
      const long long t1 = B200SDR_PROF(prm) ? clock64() : 0;

      float2 acc[RPT][MP];
      const unsigned char* block = tiles + slot * slotStride + warp * OTW * D * ES;
      if constexpr (MMA) {
        tileMmaPartials<MP, RPT, MIX>(block, bFrag, prm.kSteps, D, M, rot, prm.digitScale[0], prm.digitScale[1], prm.digitScale[2], part, lane);
      } else {
        tilePartialSums<ELEM, MIX, MP, RPT, CONV>(block, hT, W, D, lane, 32u, acc);
      }
      const long long t2 = B200SDR_PROF(prm) ? clock64() : 0;
      // ---- the slot is free once every compute warp is past its main loop; the last one refills it ----------------
      __syncwarp();
      if (t + S < nTiles) {
        // every value this warp loaded from the slot has been consumed by arithmetic, so its reads are complete
        unsigned last = 0;
        if (lane == 0) last = atomicAdd(&slotDone[slot], 1u) == NW - 1 ? 1u : 0u;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          fillTileTail(t + S, lane, 32u);
          __syncwarp();
          if (lane == 0) {
            slotDone[slot] = 0;
            fenceProxyAsync();
            issueTileInto(t + S, slot);
          }
        }
      }

      // ---- y[k] = P[k][0] + sum_{m>=1} rot[m] * P[k+m][m], inside the warp's own row block ----------------------
      if constexpr (MMA) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < RPT; i++) {
          const unsigned row = lane + 32u * i;
          float2 y = part[row];
#pragma unroll
          for (int m = 1; m < MP; m++) {
            if (m < M && row + m < RB) {
              const float2 v = part[m * RB + row + m];
              y.x += v.x;
              y.y += v.y;
            }
          }
          acc[i][0] = y;
        }
        __syncwarp();  // the scratch is rewritten by the next tile's epilogue (and by the FM line below)
      } else if (MP > 1 && M > 1) {
#pragma unroll
        for (int i = 0; i < RPT; i++) {
          const unsigned row = lane + 32u * i;
#pragma unroll
          for (int m = 1; m < MP; m++) {
            if (m < M) {
              float2 v = acc[i][m];
              if constexpr (MIX) v = cmulf(v, rot[m]);
              part[(m - 1) * RB + row] = v;
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < RPT; i++) {
          const unsigned row = lane + 32u * i;
#pragma unroll
          for (int m = 1; m < MP; m++) {
            if (m < M && row + m < RB) {
              const float2 v = part[(m - 1) * RB + row + m];
              acc[i][0].x += v.x;
              acc[i][0].y += v.y;
            }
          }
        }
      }

      // ---- demodulate into the line (after the audio warp has handed it back) -----------------------------------
      const unsigned cur = t & 1u, use = t >> 1;
      const long long t3 = B200SDR_PROF(prm) ? clock64() : 0;
      if (use > 0) mbarWait(&dmEmpty[cur], (use - 1) & 1u);
      const long long t4 = B200SDR_PROF(prm) ? clock64() : 0;
      float* line = dm + cur * prm.dmCapacity + carry + warp * OTW;
      if (fm) {
#pragma unroll
        for (int i = 0; i < RPT; i++) sums[lane + 32u * i] = acc[i][0];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < RPT; i++) {
          const unsigned row = lane + 32u * i;
          if (row < OTW) {
            const float2 c = acc[i][0], n = sums[row + 1];
            const float2 d = make_float2(fmaf(n.y, c.y, n.x * c.x), fmaf(n.y, c.x, -n.x * c.y));
            const float2 r = cmulf(d, rot1);
            line[row] = prm.gain * atan2f(r.y, r.x);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < RPT; i++) {
          const unsigned row = lane + 32u * i;
          if (row < OTW) line[row] = sqrtf(fmaf(acc[i][0].x, acc[i][0].x, acc[i][0].y * acc[i][0].y));
        }
      }
      __syncwarp();
      if (lane == 0) mbarArrive(&dmFull[cur]);

      // bookkeeping identical to the audio warp's
      const unsigned nA = outputsReady(carry);
      done += nA;
      carry = carry + OT - nA * D2;
      if (++slot == S) {
        slot = 0;
        slotPhase ^= 1u;
      }
      if (B200SDR_PROF(prm)) {
        const long long t5 = clock64();
        cWait += t1 - t0;
        cMain += t2 - t1;
        cLine += t4 - t3;
        cRest += (t3 - t2) + (t5 - t4);
      }
    }
    if (B200SDR_PROF(prm) && lane == 0) {
      unsigned long long* o = prm.prof + (static_cast<unsigned long long>(blockIdx.x) * NW + warp) * 6;
      o[0] = cWait;
      o[1] = cMain;
      o[2] = cRest;
      o[3] = cLine;
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      o[4] = (static_cast<unsigned long long>(smid) << 40) | static_cast<unsigned long long>(tLoop - tEntry);  // SM id | prologue
      o[5] = clock64() - tEntry;    // whole life of the warp
    }
  } else {
    // =========================== audio warp ===========================
    const bool pairs = (D2 & 1u) == 0;
    for (unsigned t = 0; t < nTiles; t++) {
      const unsigned cur = t & 1u, use = t >> 1;
      mbarWait(&dmFull[cur], use & 1u);
      const float* line = dm + cur * prm.dmCapacity;
      const unsigned len = carry + OT;
      const unsigned nA = outputsReady(carry);
      // each lane works on outputs o and o+32 at once: two independent dot products hide the shared-memory latency
      const unsigned aLane = (warp - NW) * 32u + lane, aLanes = NA * 32u;  // this thread's index among the audio threads
      for (unsigned o = aLane; o < nA; o += 2u * aLanes) {
        const bool two = o + aLanes < nA;
        const float* xa = line + o * D2;
        const float* xb = two ? xa + aLanes * D2 : xa;
        float a0s = 0.0f, a1s = 0.0f, a2s = 0.0f, a3s = 0.0f, b0s = 0.0f, b1s = 0.0f, b2s = 0.0f, b3s = 0.0f;
        unsigned j = 0;
        if (pairs) {  // o*D2 is even: 64-bit loads of the demod line
#pragma unroll 2
          for (; j + 4 <= T2; j += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(h2 + j);
            const float2 p0 = *reinterpret_cast<const float2*>(xa + j), p1 = *reinterpret_cast<const float2*>(xa + j + 2);
            const float2 q0 = *reinterpret_cast<const float2*>(xb + j), q1 = *reinterpret_cast<const float2*>(xb + j + 2);
            a0s = fmaf(hv.x, p0.x, a0s);
            a1s = fmaf(hv.y, p0.y, a1s);
            a2s = fmaf(hv.z, p1.x, a2s);
            a3s = fmaf(hv.w, p1.y, a3s);
            b0s = fmaf(hv.x, q0.x, b0s);
            b1s = fmaf(hv.y, q0.y, b1s);
            b2s = fmaf(hv.z, q1.x, b2s);
            b3s = fmaf(hv.w, q1.y, b3s);
          }
        } else {
#pragma unroll 2
          for (; j + 4 <= T2; j += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(h2 + j);
            a0s = fmaf(hv.x, xa[j], a0s);
            a1s = fmaf(hv.y, xa[j + 1], a1s);
            a2s = fmaf(hv.z, xa[j + 2], a2s);
            a3s = fmaf(hv.w, xa[j + 3], a3s);
            b0s = fmaf(hv.x, xb[j], b0s);
            b1s = fmaf(hv.y, xb[j + 1], b1s);
            b2s = fmaf(hv.z, xb[j + 2], b2s);
            b3s = fmaf(hv.w, xb[j + 3], b3s);
          }
        }
        for (; j < T2; j++) {
          a0s = fmaf(h2[j], xa[j], a0s);
          b0s = fmaf(h2[j], xb[j], b0s);
        }
        prm.out[a0 + done + o] = (a0s + a1s) + (a2s + a3s);
        if (two) prm.out[a0 + done + o + aLanes] = (b0s + b1s) + (b2s + b3s);
      }
      const unsigned consumed = nA * D2;
      const unsigned newCarry = len - consumed;
      float* next = dm + (cur ^ 1u) * prm.dmCapacity;
      for (unsigned i = aLane; i < newCarry; i += aLanes) next[i] = line[consumed + i];
      // every audio warp is done READING this line before any of them starts the next tile, whose carry copy writes it
      if (NA > 1) {
        asm volatile("bar.sync 1, %0;" ::"r"(NA * 32u) : "memory");
      } else {
        __syncwarp();
      }
      if (lane == 0) mbarArrive(&dmEmpty[cur]);
      done += nA;
      carry = newCarry;
    }
  }
}

#endif  // __CUDACC__

}  // namespace b200sdr
