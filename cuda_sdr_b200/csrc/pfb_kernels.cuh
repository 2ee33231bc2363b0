// pfbKernel -- RF stage of the wideband channelizer when every channel lies on a raster fs/N (N a power of two <= 256)
// with one common offset f0: polyphase filter bank + N-point inverse FFT, ALL channels at once.
//
// With w_c = w0 + 2*pi*b_c/N the per-channel chain  y_c[k] = sum_j h[j] exp(j*w_c*(kD+j)) x[kD+j]  (convert, mix, FIR,
// decimate: reference Int8ToFloat.cpp:80-100, ComplexCosineSource.cpp:52-88, Multiply.cpp:132-159, Fir.cpp:229-268, once
// per channel) factors into
//     u_k[r]  = sum_q h'[r + N q] x[kD + r + N q],    h'[j] = h[j] exp(j*w0*j) / 128          (r = 0..N-1)
//     Y_k[b]  = sum_r u_k[r] exp(+2*pi*i*b*r/N)                                                 (one inverse FFT)
//     y_c[k]  = exp(j*w_c*k*D) * Y_k[b_c]
// and AM (|y|) and the FM discriminator (arg(y[k+1] conj y[k]), up to the constant rotation exp(j*w_c*D)) do not see the
// carrier.  Work per input sample drops from ~38 flop PER CHANNEL to ~60 flop for all N bins.
//
// Precision: a strong channel's rounding noise lands in every bin, so a 1e-5 bound relative to a WEAK channel's own level
// needs more than fp32 in the filter bank and the FFT; both run in fp64 (B200 has a full FP64 pipe).  int8 samples enter the
// DFMA without a conversion instruction: PRMT drops the biased byte v = x + 128 into the mantissa of 2^20, X = 2^20 + v is
// exact, and sum h'*(X - 2^20 - 128) = sum h'*X + acc0 with acc0[r] = -(2^20 + 128)(1 + i) sum_q h'[r + N q] precomputed.
//
// One persistent CTA per SM: taps, twiddles and acc0 stay in shared memory; per tile of 32 consecutive RF outputs the input
// window is staged once (each sample is used by T/D outputs).  A tile runs in 4 rounds of 8 outputs:
//   filter bank  -- all 256 threads; a thread owns two adjacent phases r and FOUR outputs, so each pair of tap loads feeds
//                   32 DFMAs in 16 independent chains (the stage is DFMA-bound, not latency-bound);
//   FFT          -- one warp per output: in-place Stockham radix-4 (registers hold a pass, one 4 KB buffer per warp);
//   demodulation -- one warp per output over the configured channels (the FM successor is the neighbouring warp's buffer)
//                   into a [channel][32] tile that is flushed with 128-byte rows.
// With FM channels the 32nd RF output of a tile is only the successor of the 31st: tiles advance by 31.
#pragma once

#include "common.cuh"

namespace b200sdr {

constexpr unsigned kPfbTileK = 32;      // RF outputs per tile
constexpr unsigned kPfbWarps = 8;       // = outputs per round
constexpr unsigned kPfbRounds = kPfbTileK / kPfbWarps;
constexpr unsigned kPfbMaxN = 256;
constexpr unsigned kPfbOutStride = kPfbTileK + 1;  // padded: lanes = channels write one column conflict-free

struct PfbParams {
  const unsigned char* in;      // interleaved int8 I,Q; 16-byte aligned
  float* out;                   // demodulated samples [channel][outStride]
  const double* tapsRe;         // re h'[j], j < Qn * N (zero beyond T1)
  const double* tapsIm;         // im h'[j]
  const double2* acc0;          // [N]
  const double2* twiddle;       // exp(+2*pi*i*t/N), t < N
  const int* bin;               // [C] FFT bin of each channel
  const int* order;             // [C] channel indices sorted by modulation: lane i of a warp handles order[i + 32*slot], so a warp
                                //     instruction demodulates 32 channels of ONE kind (no AM / FM divergence)
  const int* mod;               // [C] 0 AM, 1 FM
  const float* gain;            // [C]
  const float2* rot1;           // [C] exp(j*w_c*D)
  unsigned long long nInBytes;
  unsigned long long nOut;      // demodulated samples per channel to produce
  unsigned long long outStride;
  unsigned D1, N, Qn, C;
  int anyFm;
  int forceAm;                  // demodulate every channel as AM whatever `mod` says (the AM tail pass of a mixed AM/FM set)
};

struct PfbSmem {
  unsigned tapsReOff, tapsImOff, acc0Off, twOff, fftOff, lastOff, outOff, winOff, winBytes, total;
};

__host__ __device__ inline PfbSmem pfbSmemLayout(unsigned N, unsigned Qn, unsigned D1, unsigned C) {
  PfbSmem s;
  unsigned off = 0;
  s.tapsReOff = off;
  off += Qn * N * 8u;
  s.tapsImOff = off;
  off += Qn * N * 8u;
  s.acc0Off = off;
  off += N * 16u;
  s.twOff = off;
  off += N * 16u;
  s.fftOff = off;
  off += kPfbWarps * N * 16u;
  s.lastOff = off;
  off += 2u * C * 8u;  // double-buffered by round parity
  s.outOff = off;
  off += C * kPfbOutStride * 4u;
  off = (off + 15u) & ~15u;
  s.winOff = off;
  s.winBytes = ((kPfbTileK - 1u) * D1 + Qn * N) * 2u;
  s.winBytes = (s.winBytes + 15u) & ~15u;
  off += s.winBytes + 16u;  // the window may start up to 12 bytes into its first 16-byte line
  s.total = off;
  return s;
}

#ifdef __CUDACC__

__device__ __forceinline__ double2 cmuld(double2 a, double2 b) { return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }

// One pass of the in-place Stockham autosort FFT (inverse sign) over buf[N]: radix R = 4, or 2 for the last pass when log2 N
// is odd.  Every lane first reads the inputs of its butterflies (<= 8 complex values), the warp synchronises, then writes.
template <int R>
__device__ __forceinline__ void pfbFftPass(double2* buf, const double2* tw, unsigned N, unsigned Ns, unsigned lane) {
  constexpr int B = 8 / R;  // butterflies per lane
  const unsigned count = N / R, twStep = N / (Ns * R);
  double2 v[B][R];
#pragma unroll
  for (int b = 0; b < B; b++) {
    const unsigned j = lane + 32u * b;
    if (j < count) {
      const unsigned kk = j & (Ns - 1u);
#pragma unroll
      for (int r = 0; r < R; r++) v[b][r] = buf[j + r * count];
      if (Ns > 1u) {  // the first pass has no twiddles; later ones load w and square it (one table read per butterfly)
        const double2 w1 = tw[kk * twStep];
        v[b][1] = cmuld(v[b][1], w1);
        if constexpr (R == 4) {
          const double2 w2 = cmuld(w1, w1);
          v[b][2] = cmuld(v[b][2], w2);
          v[b][3] = cmuld(v[b][3], cmuld(w2, w1));
        }
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int b = 0; b < B; b++) {
    const unsigned j = lane + 32u * b;
    if (j < count) {
      const unsigned kk = j & (Ns - 1u);
      const unsigned j0 = (j - kk) * R + kk;
      if constexpr (R == 4) {
        const double2 t0 = make_double2(v[b][0].x + v[b][2].x, v[b][0].y + v[b][2].y);
        const double2 t1 = make_double2(v[b][0].x - v[b][2].x, v[b][0].y - v[b][2].y);
        const double2 t2 = make_double2(v[b][1].x + v[b][3].x, v[b][1].y + v[b][3].y);
        const double2 t3 = make_double2(-(v[b][1].y - v[b][3].y), v[b][1].x - v[b][3].x);  // (v1 - v3) * (+i)
        buf[j0] = make_double2(t0.x + t2.x, t0.y + t2.y);
        buf[j0 + Ns] = make_double2(t1.x + t3.x, t1.y + t3.y);
        buf[j0 + 2u * Ns] = make_double2(t0.x - t2.x, t0.y - t2.y);
        buf[j0 + 3u * Ns] = make_double2(t1.x - t3.x, t1.y - t3.y);
      } else {
        buf[j0] = make_double2(v[b][0].x + v[b][1].x, v[b][0].y + v[b][1].y);
        buf[j0 + Ns] = make_double2(v[b][0].x - v[b][1].x, v[b][0].y - v[b][1].y);
      }
    }
  }
  __syncwarp();
}

// biased byte v = x + 128 (selected by SEL from the word) -> the double 2^20 + v, exactly, with one PRMT
template <unsigned SEL>
__device__ __forceinline__ double pfbSample(unsigned w) {
  return __hiloint2double(static_cast<int>(__byte_perm(w, 0x41300000u, SEL)), 0);
}

__global__ void __launch_bounds__(kPfbWarps * 32, 1) pfbKernel(const PfbParams prm) {
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned N = prm.N, Qn = prm.Qn, D = prm.D1, C = prm.C;
  const bool fm = prm.anyFm != 0 && prm.forceAm == 0;
  const PfbSmem lay = pfbSmemLayout(N, Qn, D, C);
  double* tapsRe = reinterpret_cast<double*>(smem + lay.tapsReOff);
  double* tapsIm = reinterpret_cast<double*>(smem + lay.tapsImOff);
  double2* acc0 = reinterpret_cast<double2*>(smem + lay.acc0Off);
  double2* tw = reinterpret_cast<double2*>(smem + lay.twOff);
  double2* fftBufs = reinterpret_cast<double2*>(smem + lay.fftOff);
  float2* lastY = reinterpret_cast<float2*>(smem + lay.lastOff);
  float* outTile = reinterpret_cast<float*>(smem + lay.outOff);
  unsigned char* win = smem + lay.winOff;
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  double2* buf = fftBufs + warp * N;

  for (unsigned i = tid; i < Qn * N; i += blockDim.x) {
    tapsRe[i] = prm.tapsRe[i];
    tapsIm[i] = prm.tapsIm[i];
  }
  for (unsigned i = tid; i < N; i += blockDim.x) {
    acc0[i] = prm.acc0[i];
    tw[i] = prm.twiddle[i];
  }

  // the channels of this lane (order[lane + 32 * slot]) and their constants live in registers for the whole kernel
  constexpr int kSlots = kPfbMaxN / 32;
  int chL[kSlots], binL[kSlots], modL[kSlots];
  float gainL[kSlots];
  float2 rotL[kSlots];
#pragma unroll
  for (int slot = 0; slot < kSlots; slot++) {
    const unsigned i = lane + 32u * slot;
    const int ch = i < C ? prm.order[i] : -1;
    chL[slot] = ch;
    binL[slot] = ch >= 0 ? prm.bin[ch] : 0;
    modL[slot] = ch >= 0 && prm.forceAm == 0 ? prm.mod[ch] : 0;
    gainL[slot] = ch >= 0 ? prm.gain[ch] : 0.0f;
    rotL[slot] = ch >= 0 ? prm.rot1[ch] : make_float2(1.0f, 0.0f);
  }

  const unsigned tileOut = kPfbTileK - (fm ? 1u : 0u);  // demodulated samples a tile yields
  const unsigned long long nRf = prm.nOut + (fm ? 1ull : 0ull);  // RF outputs that exist for this call
  const unsigned long long tiles = (prm.nOut + tileOut - 1) / tileOut;

  // The input window of a tile is staged in four chunks, one per round, each fetched into registers while the previous
  // round's filter bank runs and stored after that stage's barrier (the filter bank is the only reader of the window, and
  // chunk rho + 1 lies past everything round rho reads): global-memory latency hides behind the DFMAs.
  constexpr int kPre = 6;  // 16-byte lines per thread held in flight
  uint4 pre[kPre];
  unsigned preLo = 0, preHi = 0;            // lines [preLo, preHi) of the window they belong to
  unsigned long long preFirst = 0;          // byte address of line 0 of that window
  auto windowStart = [&](unsigned long long t) { return t * tileOut * D * 2ull; };
  auto chunkEnd = [&](unsigned skew, unsigned rho) {  // lines of the window that rounds 0..rho read
    return (skew + ((kPfbWarps * rho + kPfbWarps - 1u) * D + Qn * N) * 2u + 15u) / 16u;
  };
  auto fetchLine = [&](unsigned long long b) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (b + 16ull <= prm.nInBytes) {
      v = ldStream(reinterpret_cast<const uint4*>(prm.in + b));
    } else if (b < prm.nInBytes) {  // the last, partial line of the input
      unsigned char tmp[16];
      for (unsigned e = 0; e < 16u; e++) tmp[e] = b + e < prm.nInBytes ? prm.in[b + e] : 0;
      v = *reinterpret_cast<const uint4*>(tmp);
    }
    return v;
  };
  auto loadChunk = [&](unsigned long long t, unsigned rho) {
    const unsigned long long start = windowStart(t);
    const unsigned skew = static_cast<unsigned>(start & 15ull);
    preFirst = start - skew;
    preLo = rho == 0 ? 0u : chunkEnd(skew, rho - 1u);
    preHi = chunkEnd(skew, rho);
#pragma unroll
    for (int i = 0; i < kPre; i++) {
      const unsigned line = preLo + tid + static_cast<unsigned>(i) * blockDim.x;
      if (line < preHi) pre[i] = fetchLine(preFirst + 16ull * line);
    }
  };
  auto storeChunk = [&]() {
    uint4* dst = reinterpret_cast<uint4*>(win);
    auto put = [&](unsigned line, uint4 v) {
      v.x ^= 0x80808080u;  // bias every byte: v = x + 128
      v.y ^= 0x80808080u;
      v.z ^= 0x80808080u;
      v.w ^= 0x80808080u;
      dst[line] = v;
    };
#pragma unroll
    for (int i = 0; i < kPre; i++) {
      const unsigned line = preLo + tid + static_cast<unsigned>(i) * blockDim.x;
      if (line < preHi) put(line, pre[i]);
    }
    for (unsigned line = preLo + tid + kPre * blockDim.x; line < preHi; line += blockDim.x) put(line, fetchLine(preFirst + 16ull * line));  // long rows only
  };
  if (blockIdx.x < tiles) {
    loadChunk(blockIdx.x, 0);
    storeChunk();
  }

  for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const unsigned long long k0 = tile * tileOut;
    const unsigned winSkew = static_cast<unsigned>(windowStart(tile) & 15ull);  // the window starts this far into its first line
    __syncthreads();  // chunk 0 of the window and the tables are in place; the previous tile's output tile has been flushed

    for (unsigned round = 0; round < kPfbRounds; round++) {
      preLo = preHi = 0;
      if (round + 1u < kPfbRounds) {
        loadChunk(tile, round + 1u);
      } else if (tile + gridDim.x < tiles) {
        loadChunk(tile + gridDim.x, 0);
      }
      // ---- filter bank: u[r] of the round's 8 outputs; thread = (pair of phases, group of 4 outputs) ----
      for (unsigned item = tid; item < N; item += blockDim.x) {
        const unsigned pair = item % (N / 2u), kg = item / (N / 2u);
        const unsigned r = 2u * pair;
        const unsigned kk0 = round * kPfbWarps + kg * 4u;  // first of this thread's four outputs (within the tile)
        const double2 a0 = acc0[r], a1 = acc0[r + 1u];
        double2 u[4][2];
#pragma unroll
        for (int o = 0; o < 4; o++) {
          u[o][0] = a0;
          u[o][1] = a1;
        }
        const unsigned char* x0 = win + winSkew + (static_cast<size_t>(kk0) * D + r) * 2u;
#pragma unroll 2
        for (unsigned q = 0; q < Qn; q++) {
          const double2 hr = *reinterpret_cast<const double2*>(tapsRe + q * N + r);  // re h'[r], re h'[r + 1]
          const double2 hi = *reinterpret_cast<const double2*>(tapsIm + q * N + r);
#pragma unroll
          for (int o = 0; o < 4; o++) {
            const unsigned w = *reinterpret_cast<const unsigned*>(x0 + (static_cast<size_t>(o) * D + q * N) * 2u);  // I0 Q0 I1 Q1, biased
            const double i0 = pfbSample<0x7650>(w), q0 = pfbSample<0x7651>(w), i1 = pfbSample<0x7652>(w), q1 = pfbSample<0x7653>(w);
            u[o][0].x = fma(hr.x, i0, fma(-hi.x, q0, u[o][0].x));
            u[o][0].y = fma(hr.x, q0, fma(hi.x, i0, u[o][0].y));
            u[o][1].x = fma(hr.y, i1, fma(-hi.y, q1, u[o][1].x));
            u[o][1].y = fma(hr.y, q1, fma(hi.y, i1, u[o][1].y));
          }
        }
#pragma unroll
        for (int o = 0; o < 4; o++) {
          double2* dst = fftBufs + (kg * 4u + o) * N + r;
          dst[0] = u[o][0];
          dst[1] = u[o][1];
        }
      }
      __syncthreads();
      storeChunk();  // the window's next chunk (this round's filter bank was the last reader of anything it overwrites)
      // ---- inverse FFT of output kk = round * 8 + warp: natural order in and out ----
      const unsigned kk = round * kPfbWarps + warp;
      const unsigned long long k = k0 + kk;
      if (k < nRf) {  // warp-uniform
        unsigned Ns = 1;
        for (; Ns * 4u <= N; Ns *= 4u) pfbFftPass<4>(buf, tw, N, Ns, lane);
        if (Ns < N) pfbFftPass<2>(buf, tw, N, Ns, lane);
      }
      __syncthreads();
      // ---- demodulate the configured channels; FM pairs this output with its predecessor (the neighbouring warp's) ----
      if (k < nRf) {
        const double2* prevBuf = warp > 0 ? buf - N : nullptr;
#pragma unroll
        for (int slot = 0; slot < kSlots; slot++) {
          if (chL[slot] >= 0) {
            const unsigned ch = static_cast<unsigned>(chL[slot]);
            const double2 yd = buf[binL[slot]];
            if (modL[slot] == 0) {
              {
                const float yr = static_cast<float>(yd.x), yi = static_cast<float>(yd.y);
                outTile[ch * kPfbOutStride + kk] = sqrtf(fmaf(yr, yr, yi * yi));
              }
            } else if (kk > 0) {
              float2 c;
              if (prevBuf) {
                const double2 pd = prevBuf[binL[slot]];
                c = make_float2(static_cast<float>(pd.x), static_cast<float>(pd.y));
              } else {
                c = lastY[(round & 1u) * C + ch];  // the last output of the previous round
              }
              const float2 y = make_float2(static_cast<float>(yd.x), static_cast<float>(yd.y));
              const float2 d = make_float2(fmaf(y.y, c.y, y.x * c.x), fmaf(y.y, c.x, -y.x * c.y));
              const float2 e = make_float2(fmaf(d.x, rotL[slot].x, -d.y * rotL[slot].y), fmaf(d.x, rotL[slot].y, d.y * rotL[slot].x));
              outTile[ch * kPfbOutStride + kk - 1u] = gainL[slot] * atan2f(e.y, e.x);
            }
          }
        }
      }
      if (fm && warp == kPfbWarps - 1u && k < nRf) {  // this round's last output is the next round's first predecessor
#pragma unroll
        for (int slot = 0; slot < kSlots; slot++) {
          if (chL[slot] >= 0) {
            const unsigned ch = static_cast<unsigned>(chL[slot]);
            const double2 yd = buf[binL[slot]];
            lastY[((round + 1u) & 1u) * C + ch] = make_float2(static_cast<float>(yd.x), static_cast<float>(yd.y));
          }
        }
      }
      __syncthreads();  // the buffers are free for the next round's filter bank
    }
    __syncthreads();
    // ---- flush the [channel][tileOut] tile: one row per channel ----
    for (unsigned i = tid; i < C * kPfbTileK; i += blockDim.x) {
      const unsigned ch = i / kPfbTileK, kk = i % kPfbTileK;
      if (kk < tileOut && k0 + kk < prm.nOut) prm.out[static_cast<unsigned long long>(ch) * prm.outStride + k0 + kk] = outTile[ch * kPfbOutStride + kk];
    }
  }
}

#endif  // __CUDACC__

}  // namespace b200sdr
