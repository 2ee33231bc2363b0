// pfb256Kernel -- the polyphase-filter-bank + FFT channelizer (see pfb_kernels.cuh for the algebra) specialised for the raster
// N = 256 (BASELINE config C5: 256 channels on fs/256), built around what the first kernel was measured to be bound by
// (profiles/r1g_c5_summary.md: 8 warps per SM, every phase behind a CTA-wide barrier, taps re-read from shared memory by every
// thread, four Stockham passes through shared memory with 31 % bank-conflict wavefronts):
//
//   * TAPS IN REGISTERS.  A thread owns two adjacent phases r, r+1 for the whole kernel; its QN complex taps per phase
//     (fp64, 4*QN doubles) never leave the register file, so the filter bank's only shared-memory traffic is one 32-bit
//     sample load per eight DFMAs, and 70 KB of shared memory are free.
//   * TWO CTAs PER SM of 128 threads (4 warps), each with its own contiguous run of RF outputs.  While one CTA sits in the
//     latency-bound FFT / demodulation of a round, the other one's filter bank keeps the FP64 pipe busy; there is no
//     producer / consumer hand-off between them, only the hardware's warp scheduler.
//   * THE INPUT STREAM IS A TMA RING.  A CTA walks its outputs in rounds of four; each round advances the input by 4*D
//     samples, and an elected thread keeps three rounds of chunks in flight with cp.async.bulk on per-chunk mbarriers
//     (32 KB ring, no register staging, no barrier for the loads).
//   * REGISTER-RESIDENT FFT.  One warp per RF output, 8 points per lane: 256 = 8 x 8 x 4, two 8-point butterflies and two
//     4-point butterflies per lane in registers, two transposes through the warp's own 4 KB buffer with padded, conflict-free
//     layouts (33-element rows; the filter bank stores phase r at position r/2 + 128*(r&1) so that its own stores and the
//     FFT's first loads are conflict-free as well).  Shared-memory traffic per FFT: 20 KB instead of 36 KB, no conflicts.
//   * Demodulated samples leave through a [channel][32] tile flushed as 128-byte rows every eight rounds.
//
// Arithmetic is the first kernel's: fp64 filter bank (int8 -> double by one PRMT, see pfbSample) and fp64 FFT, float
// demodulation.  Every RF output is a function of its own input window only, so results do not depend on how outputs are
// split over CTAs, launches or GPUs (time segments concatenate bit-exactly).
#pragma once

#include <type_traits>

#include "pfb_kernels.cuh"

namespace b200sdr {

constexpr unsigned kP2Threads = 128;
constexpr unsigned kP2Ring = 32768;       // input ring, bytes (power of two)
constexpr unsigned kP2Prefetch = 3;       // chunks (rounds) in flight beyond the ones a round reads
constexpr unsigned kP2Bars = 16;          // mbarriers cycled over the chunks
constexpr unsigned kP2FftStride = 264;    // complex doubles per FFT buffer: 8 rows of 33
constexpr unsigned kP2OutStride = 33;     // floats per channel row of the output tile

struct Pfb256Params {
  const unsigned char* in;      // interleaved int8 I,Q; 16-byte aligned
  float* out;                   // demodulated samples [channel][outStride]
  const double* tapsRe;         // re h'[j], j < Qn * 256
  const double* tapsIm;
  const double2* acc0;          // [256]
  const double2* twiddle;       // exp(+2*pi*i*t/256), t < 256
  const float4* chanInfo;       // [256] in demodulation order (sorted by modulation): gain, rot1.x, rot1.y, bits
                                //   bits = channel | bin << 16 | fm << 24 | valid << 25
  unsigned long long nInBytes;
  unsigned long long nOut;      // demodulated samples per channel to produce
  unsigned long long outStride;
  unsigned D1, Qn, C;
  int anyFm;                    // 0: no FM channel (or every channel forced to AM)
  int forceAm;
};

struct Pfb256Smem {
  unsigned barOff, infoOff, twOff, fftOff, yOff, outOff, ringOff, total;
};
__host__ __device__ inline Pfb256Smem pfb256SmemLayout() {
  Pfb256Smem s;
  unsigned off = 0;
  s.barOff = off;
  off += 128;
  s.infoOff = off;
  off += 256 * 16;
  s.twOff = off;
  off += 256 * 16;
  s.fftOff = off;
  off += 4 * kP2FftStride * 16;
  s.yOff = off;
  off += 2 * 4 * 256 * 8;
  s.outOff = off;
  off += 256 * kP2OutStride * 4;
  off = (off + 127u) & ~127u;
  s.ringOff = off;
  off += kP2Ring;
  s.total = off;
  return s;
}

// chunks (= rounds) whose bytes a round reads: ceil((6 D + QN * 512) / (8 D))
__host__ __device__ inline unsigned pfb256LiveChunks(unsigned D1, unsigned QN) { return (6u * D1 + QN * 512u + 8u * D1 - 1u) / (8u * D1); }
__host__ __device__ inline bool pfb256Fits(unsigned D1, unsigned QN) {
  const unsigned live = pfb256LiveChunks(D1, QN);
  return D1 % 8u == 0 && (live + kP2Prefetch) * 8u * D1 <= kP2Ring && live + kP2Prefetch < kP2Bars;
}

#ifdef __CUDACC__

__device__ __forceinline__ void p2FenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmulI(double2 a) { return make_double2(-a.y, a.x); }  // a * i

// atan2 for the FM discriminator: z = min/max in [0, 1], odd minimax polynomial of degree 15 (max error 1.4e-7 rad in float,
// fitted on [0, 1]; coefficients derived in tools/atan_fit.py), quadrant by symmetry.  About half the instructions of atan2f.
__device__ __forceinline__ float p2Atan2(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float z = mx > 0.0f ? __fdividef(mn, mx) : 0.0f;
  const float t = z * z;
  float p = -0.004054363816976547f;
  p = fmaf(p, t, 0.021862192079424858f);
  p = fmaf(p, t, -0.055911172181367874f);
  p = fmaf(p, t, 0.09642109274864197f);
  p = fmaf(p, t, -0.13908593356609344f);
  p = fmaf(p, t, 0.1994655877351761f);
  p = fmaf(p, t, -0.33329859375953674f);
  p = fmaf(p, t, 0.9999993443489075f);
  float r = p * z;
  if (ay > ax) r = 1.57079632679489661923f - r;
  if (x < 0.0f) r = 3.14159265358979323846f - r;
  return y < 0.0f ? -r : r;
}

// 4-point inverse DFT: c[k] <- sum_n c[n] i^(n k)
__device__ __forceinline__ void dft4inv(double2& c0, double2& c1, double2& c2, double2& c3) {
  const double2 s0 = cadd(c0, c2), s1 = csub(c0, c2), s2 = cadd(c1, c3), s3 = cmulI(csub(c1, c3));
  c0 = cadd(s0, s2);
  c1 = cadd(s1, s3);
  c2 = csub(s0, s2);
  c3 = csub(s1, s3);
}

// 8-point inverse DFT in place: v[b] <- sum_n v[n] exp(+2 pi i n b / 8)   (radix 2: even bins from a, odd bins from b)
__device__ __forceinline__ void dft8inv(double2 (&v)[8]) {
  constexpr double kH = 0.70710678118654752440;
  double2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  double2 b0 = csub(v[0], v[4]), t1 = csub(v[1], v[5]), t2 = csub(v[2], v[6]), t3 = csub(v[3], v[7]);
  const double2 b1 = make_double2((t1.x - t1.y) * kH, (t1.x + t1.y) * kH);    // * exp(+i pi/4)
  const double2 b2 = cmulI(t2);                                               // * i
  const double2 b3 = make_double2((-t3.x - t3.y) * kH, (t3.x - t3.y) * kH);   // * exp(+3 i pi/4)
  dft4inv(a0, a1, a2, a3);
  double2 c0 = b0, c1 = b1, c2 = b2, c3 = b3;
  dft4inv(c0, c1, c2, c3);
  v[0] = a0;
  v[2] = a1;
  v[4] = a2;
  v[6] = a3;
  v[1] = c0;
  v[3] = c1;
  v[5] = c2;
  v[7] = c3;
}

// One 256-point inverse FFT by one warp, in place in `buf` (kP2FftStride complex doubles; on entry phase r sits at position
// r/2 + 128*(r&1)).  On return y[m] = Y[lane + 32 m].
// twA[(b0 - 1) * 32 + lane] = W256^(b0 * perm(lane)) and twB[(c0 - 1) * 4 + s0] = W32^(c0 * s0): the twiddles in the order the
// lanes read them (consecutive lanes, consecutive 16-byte entries: no bank conflicts; a gather from the plain 256-entry
// table costs up to 8 wavefronts per load).
__device__ __forceinline__ void pfb256Fft(double2* buf, const double2* twA, const double2* twB, unsigned lane, double2 (&y)[8]) {
  // step 1: lane <-> r0 = perm(lane) (even phases on lanes 0..15, odd ones on 16..31); 8-point DFT over r1, r = r0 + 32 r1
  const unsigned pos0 = (lane & 15u) + (lane >= 16u ? 128u : 0u);
  double2 v[8];
#pragma unroll
  for (int r1 = 0; r1 < 8; r1++) v[r1] = buf[pos0 + 16u * r1];
  dft8inv(v);  // v[b0] = A[b0][r0]
#pragma unroll
  for (int b0 = 1; b0 < 8; b0++) v[b0] = cmuld(v[b0], twA[(b0 - 1) * 32 + lane]);  // * W256^(b0 r0)
  __syncwarp();  // every lane has read its inputs: the buffer may be overwritten
#pragma unroll
  for (int b0 = 0; b0 < 8; b0++) buf[b0 * 33u + lane] = v[b0];  // transpose 1: row b0, column = lane (conflict-free)
  __syncwarp();
  // step 2a: lane <-> (b0, s0) = (lane & 7, lane >> 3); 8-point DFT over s1, r0 = 4 s1 + s0
  const unsigned b0 = lane & 7u, s0 = lane >> 3;
#pragma unroll
  for (int s1 = 0; s1 < 8; s1++) {
    const unsigned r = 4u * s1 + s0;                       // the r0 whose A' is wanted; it was written by lane perm^-1(r)
    v[s1] = buf[b0 * 33u + (r >> 1) + ((r & 1u) << 4)];
  }
  dft8inv(v);  // v[c0] = B[b0][c0][s0]
#pragma unroll
  for (int c0 = 1; c0 < 8; c0++) v[c0] = cmuld(v[c0], twB[(c0 - 1) * 4 + s0]);  // * W32^(c0 s0)
  __syncwarp();
#pragma unroll
  for (int c0 = 0; c0 < 8; c0++) buf[s0 * 64u + b0 + 8u * c0] = v[c0];  // transpose 2: [s0][b0 + 8 c0]
  __syncwarp();
  // step 2b: items I = lane and lane + 32 (I = b0 + 8 c0); 4-point DFT over s0: Y[I + 64 c1]
  double2 p[4], q[4];
#pragma unroll
  for (int s = 0; s < 4; s++) {
    p[s] = buf[s * 64u + lane];
    q[s] = buf[s * 64u + 32u + lane];
  }
  dft4inv(p[0], p[1], p[2], p[3]);
  dft4inv(q[0], q[1], q[2], q[3]);
#pragma unroll
  for (int c1 = 0; c1 < 4; c1++) {
    y[2 * c1] = p[c1];      // bin lane + 64 c1       = lane + 32 (2 c1)
    y[2 * c1 + 1] = q[c1];  // bin lane + 32 + 64 c1  = lane + 32 (2 c1 + 1)
  }
  __syncwarp();  // the buffer is free for the next round's filter bank once the CTA has synchronised
}

template <int QN>
__global__ void __launch_bounds__(kP2Threads, 2) pfb256Kernel(const Pfb256Params prm) {
  extern __shared__ __align__(128) unsigned char smem[];
  const Pfb256Smem lay = pfb256SmemLayout();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.barOff);
  const float4* chanInfo = reinterpret_cast<const float4*>(smem + lay.infoOff);
  double2* twA = reinterpret_cast<double2*>(smem + lay.twOff);  // 7 x 32 entries
  double2* twB = twA + 7 * 32;                                  // 7 x 4 entries
  double2* fftBufs = reinterpret_cast<double2*>(smem + lay.fftOff);
  float2* yBufs = reinterpret_cast<float2*>(smem + lay.yOff);  // [parity][warp][256]
  float* outTile = reinterpret_cast<float*>(smem + lay.outOff);
  unsigned char* ring = smem + lay.ringOff;
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const unsigned D = prm.D1, C = prm.C;
  const bool fm = prm.anyFm != 0 && prm.forceAm == 0;

  // ---- this CTA's run of demodulated samples: a multiple of 8 (32-byte sectors of the output rows, 128-byte input alignment) ---
  const unsigned long long units = (prm.nOut + 7ull) / 8ull;
  const unsigned long long perCta = units / gridDim.x, extraCta = units % gridDim.x;
  const unsigned long long unit0 = blockIdx.x * perCta + (blockIdx.x < extraCta ? blockIdx.x : extraCta);
  const unsigned long long unit1 = unit0 + perCta + (blockIdx.x < extraCta ? 1ull : 0ull);
  if (unit1 == unit0) return;
  const unsigned long long dBeg = unit0 * 8ull, dEnd = unit1 * 8ull < prm.nOut ? unit1 * 8ull : prm.nOut;
  const unsigned nK = static_cast<unsigned>(dEnd - dBeg) + (fm ? 1u : 0u);  // RF outputs dBeg .. dBeg + nK - 1
  const unsigned rounds = (nK + 3u) / 4u;
  const unsigned long long s0Bytes = dBeg * D * 2ull;                        // stream position of round 0
  const unsigned CH = 8u * D;                                                // bytes the input advances per round
  const unsigned live = pfb256LiveChunks(D, QN);

  // ---- tables ---------------------------------------------------------------------------------------------------------
  for (unsigned i = tid; i < 256u; i += kP2Threads) {
    reinterpret_cast<float4*>(smem + lay.infoOff)[i] = i < C ? prm.chanInfo[i] : make_float4(0.0f, 1.0f, 0.0f, 0.0f);
    if (i < 7u * 32u) {
      const unsigned b0 = i / 32u + 1u, l = i & 31u, r0 = l < 16u ? 2u * l : 2u * (l - 16u) + 1u;
      twA[i] = prm.twiddle[(b0 * r0) & 255u];
    } else if (i < 7u * 32u + 28u) {
      const unsigned e = i - 7u * 32u, c0 = e / 4u + 1u, s0 = e & 3u;
      twB[e] = prm.twiddle[(8u * c0 * s0) & 255u];
    }
  }
  if (tid == 0) {
    for (unsigned i = 0; i < kP2Bars; i++) mbarInit(&bars[i], 1);
    fenceMbarInit();
  }
  // taps of phases r = 2 tid, 2 tid + 1 (registers, for the whole kernel) and the start values of the accumulators
  const unsigned r = 2u * tid;
  double2 hr[QN], hi[QN];
#pragma unroll
  for (int q = 0; q < QN; q++) {
    const bool have = static_cast<unsigned>(q) < prm.Qn;
    hr[q] = have ? *reinterpret_cast<const double2*>(prm.tapsRe + q * 256 + r) : make_double2(0.0, 0.0);
    hi[q] = have ? *reinterpret_cast<const double2*>(prm.tapsIm + q * 256 + r) : make_double2(0.0, 0.0);
  }
  const double2 acc0a = prm.acc0[r], acc0b = prm.acc0[r + 1u];
  // The last tap of a phase exists only for the first T - 256 (QN - 1) phases (T = 4097: phase 0 alone); a warp whose 64 phases all
  // have a zero there skips that step of the filter bank -- adding 0 * x changes nothing, so the result is the same bit for bit.
  const bool lastTapLive =
      __any_sync(0xffffffffu, hr[QN - 1].x != 0.0 || hr[QN - 1].y != 0.0 || hi[QN - 1].x != 0.0 || hi[QN - 1].y != 0.0) != 0;
  __syncthreads();

  // (thread 0) chunk c = stream bytes [c CH, (c + 1) CH) relative to s0Bytes -> ring, clipped to the input
  auto issueChunk = [&](unsigned c) {
    uint64_t* bar = &bars[c % kP2Bars];
    const unsigned long long start = s0Bytes + static_cast<unsigned long long>(c) * CH;
    unsigned len = 0;
    if (start < prm.nInBytes) len = prm.nInBytes - start < CH ? static_cast<unsigned>(prm.nInBytes - start) : CH;
    const unsigned bulk = len & ~15u;
    const unsigned ringPos = (c * CH) & (kP2Ring - 1u);
    for (unsigned b = bulk; b < len; b++) ring[(ringPos + b) & (kP2Ring - 1u)] = prm.in[start + b];  // the input's last, partial 16 bytes
    if (bulk == 0) {
      mbarArrive(bar);
      return;
    }
    p2FenceProxyAsync();  // generic-proxy reads of the ring bytes being replaced come before the async-proxy write
    mbarExpectTx(bar, bulk);
    const unsigned first = bulk < kP2Ring - ringPos ? bulk : kP2Ring - ringPos;  // a chunk may wrap around the ring's end
    tmaBulkLoad(ring + ringPos, prm.in + start, first, bar);
    if (first < bulk) tmaBulkLoad(ring, prm.in + start + first, bulk - first, bar);
  };
  if (tid == 0)
    for (unsigned c = 0; c < live + kP2Prefetch; c++) issueChunk(c);

  unsigned flushed = 0;  // blocks of 32 written out so far
  auto flushBlock = [&](unsigned block) {  // all threads; outTile holds columns of block `block`
    const unsigned long long base = dBeg + 32ull * block;
    const unsigned valid = dEnd - base < 32ull ? static_cast<unsigned>(dEnd - base) : 32u;
    for (unsigned slot = warp; slot < C; slot += kP2Threads / 32u) {  // one 128-byte row per warp and pass
      const unsigned ch = __float_as_uint(chanInfo[slot].w) & 0xffffu;
      if (lane < valid) prm.out[static_cast<unsigned long long>(ch) * prm.outStride + base + lane] = outTile[slot * kP2OutStride + lane];
    }
  };

  for (unsigned j = 0; j < rounds; j++) {
    // ---- the chunks this round reads have landed (earlier ones were waited for by earlier rounds) ----
    {
      const unsigned c = j + live - 1u;
      mbarWait(&bars[c % kP2Bars], (c / kP2Bars) & 1u);
      if (j == 0)
        for (unsigned cc = 0; cc + 1u < live; cc++) mbarWait(&bars[cc % kP2Bars], (cc / kP2Bars) & 1u);
    }
    // ---- filter bank: u[r], u[r+1] of the round's four outputs ----
    double2 ua[4], ub[4];
#pragma unroll
    for (int o = 0; o < 4; o++) {
      ua[o] = acc0a;
      ub[o] = acc0b;
    }
    const unsigned roundBase = (j * CH + 2u * r) & (kP2Ring - 1u);
    // The round's window (four outputs 2 D bytes apart, QN taps 512 bytes apart) wraps around the ring's end in about a third of
    // the rounds; in the others every load is base-of-output + constant, with no address arithmetic per sample.
    const bool wraps = roundBase + 6u * D + static_cast<unsigned>(QN - 1) * 512u + 4u > kP2Ring;
    const unsigned char* outBase[4];
#pragma unroll
    for (int o = 0; o < 4; o++) outBase[o] = ring + roundBase + static_cast<unsigned>(o) * 2u * D;
    auto tapStep = [&](int q, auto wrapTag) {
#pragma unroll
      for (int o = 0; o < 4; o++) {
        unsigned w;
        if constexpr (decltype(wrapTag)::value) {
          const unsigned off = (roundBase + static_cast<unsigned>(o) * 2u * D + static_cast<unsigned>(q) * 512u) & (kP2Ring - 1u);
          w = *reinterpret_cast<const unsigned*>(ring + off);
        } else {
          w = *reinterpret_cast<const unsigned*>(outBase[o] + q * 512);
        }
        w ^= 0x80808080u;  // I0 Q0 I1 Q1, biased to unsigned
        const double i0 = pfbSample<0x7650>(w), q0 = pfbSample<0x7651>(w), i1 = pfbSample<0x7652>(w), q1 = pfbSample<0x7653>(w);
        ua[o].x = fma(hr[q].x, i0, fma(-hi[q].x, q0, ua[o].x));
        ua[o].y = fma(hr[q].x, q0, fma(hi[q].x, i0, ua[o].y));
        ub[o].x = fma(hr[q].y, i1, fma(-hi[q].y, q1, ub[o].x));
        ub[o].y = fma(hr[q].y, q1, fma(hi[q].y, i1, ub[o].y));
      }
    };
    if (wraps) {
#pragma unroll
      for (int q = 0; q < QN - 1; q++) tapStep(q, std::true_type());
      if (lastTapLive) tapStep(QN - 1, std::true_type());
    } else {
#pragma unroll
      for (int q = 0; q < QN - 1; q++) tapStep(q, std::false_type());
      if (lastTapLive) tapStep(QN - 1, std::false_type());
    }
#pragma unroll
    for (int o = 0; o < 4; o++) {
      double2* dst = fftBufs + o * kP2FftStride;
      dst[tid] = ua[o];          // phase r     (even) -> position r / 2
      dst[128u + tid] = ub[o];   // phase r + 1 (odd)  -> position 128 + r / 2
    }
    __syncthreads();  // (A) the four u vectors are complete; chunk j of the ring is dead
    if (tid == 0) issueChunk(j + live + kP2Prefetch);

    // ---- inverse FFT of RF output k = dBeg + 4 j + warp; Y as float2 for the demodulators ----
    const unsigned kRel = 4u * j + warp;  // relative to dBeg
    float2* yMine = yBufs + ((j & 1u) * 4u + warp) * 256u;
    {
      double2 y[8];
      pfb256Fft(fftBufs + warp * kP2FftStride, twA, twB, lane, y);
#pragma unroll
      for (int m = 0; m < 8; m++) yMine[lane + 32u * m] = make_float2(static_cast<float>(y[m].x), static_cast<float>(y[m].y));
    }
    __syncthreads();  // (B) every warp's Y is visible (the FM discriminator pairs neighbouring outputs)

    // ---- demodulate the round's four RF outputs: thread <-> channel slots tid and tid + 128 (slots are sorted by modulation, so
    // a warp runs ONE kind per pass: in C5 every thread owns one AM and one FM channel).  AM sample k comes from Y[k]; FM sample
    // k - 1 pairs Y[k - 1] (the previous output; for the round's first one the previous round's last) with Y[k].
    const float2* yRound = yBufs + (j & 1u) * 4u * 256u;
    const float2* yLast = yBufs + (((j & 1u) ^ 1u) * 4u + 3u) * 256u;
    const unsigned nValid = static_cast<unsigned>(dEnd - dBeg);
    auto demodulate = [&](bool firstFmOnly, bool skipFirstFm) {
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const unsigned slot = tid + 128u * half;
        if (slot >= C) continue;
        const float4 info = chanInfo[slot];
        const unsigned bits = __float_as_uint(info.w);
        const unsigned bin = (bits >> 16) & 0xffu;
        const bool isFm = ((bits >> 24) & 1u) != 0 && fm;
        float* row = outTile + slot * kP2OutStride;
        if (!isFm) {
          if (firstFmOnly) continue;
#pragma unroll
          for (int o = 0; o < 4; o++) {
            const unsigned k = 4u * j + o;
            const float2 y = yRound[o * 256 + bin];
            const float p = fmaf(y.x, y.x, y.y * y.y);
            if (k < nValid) row[k & 31u] = p > 0.0f ? p * rsqrtf(p) : 0.0f;
          }
        } else {
          float2 c = yLast[bin];
#pragma unroll
          for (int o = 0; o < 4; o++) {
            const unsigned k = 4u * j + o;
            const float2 y = yRound[o * 256 + bin];
            const bool wanted = o == 0 ? !skipFirstFm : !firstFmOnly;
            if (wanted && k >= 1u && k - 1u < nValid && k < nK) {
              const float2 d = make_float2(fmaf(y.y, c.y, y.x * c.x), fmaf(y.y, c.x, -y.x * c.y));
              const float2 e = make_float2(fmaf(d.x, info.y, -d.y * info.z), fmaf(d.x, info.z, d.y * info.y));
              row[(k - 1u) & 31u] = info.x * p2Atan2(e.y, e.x);
            }
            c = y;
          }
        }
      }
    };
    if (fm) {
      // FM samples lag one RF output: block b is complete once round 8 (b + 1) has produced FM sample 32 b + 31
      if (j > 0 && (j & 7u) == 0) {
        demodulate(true, false);
        __syncthreads();
        flushBlock(flushed++);
        __syncthreads();
        demodulate(false, true);
      } else {
        demodulate(false, false);
      }
    } else {
      demodulate(false, false);
      if ((j & 7u) == 7u) {
        __syncthreads();
        flushBlock(flushed++);  // the next round's demodulation starts behind its own barriers (A), (B)
      }
    }
  }
  __syncthreads();
  if (dBeg + 32ull * flushed < dEnd) flushBlock(flushed);
}

#endif  // __CUDACC__

}  // namespace b200sdr
