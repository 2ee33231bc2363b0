// Decimating-FIR kernels of the hot path (sm_100a).
//
//   rowsKernel   -- the fast path.  One thread owns RPT "rows" of D consecutive input samples (one
//                   decimation period each).  With taps re-laid out as hT[p][m] = h[m*D + p] a row q
//                   contributes the M = ceil(T/D) partial sums
//                       P[q][m] = sum_{p<D} hT[p][m] * z[q*D + p],      y[k] = sum_{m<M} P[k+m][m]
//                   so every input sample is converted and mixed exactly ONCE and only kept outputs are
//                   computed (polyphase decimation).  The raw tile (rows x D samples, int8 or float) is
//                   staged in shared memory by one 1-D TMA bulk copy; taps, the per-phase mixer table
//                   W[p] = exp(j*w*p) and the per-partial rotations exp(j*w*m*D) sit in shared memory and
//                   are read as warp-wide broadcasts; the partial sums are exchanged through shared
//                   memory once per tile.  The per-row carrier phase exp(j*w*q*D) never has to be
//                   evaluated for AM/FM outputs (|.| and the discriminator are invariant to it).
//   directKernel -- generic fallback for any (T, D), alignment and tap type: one thread per output,
//                   taps (optionally pre-multiplied by the mixer phasor) chunked through shared memory.
#pragma once

#include "common.cuh"

namespace b200sdr {

enum ElemKind : int { kElemInt8Complex = 0, kElemComplex = 1, kElemReal = 2 };
enum ModKind : int { kModAm = 0, kModFm = 1, kModNone = 2 };

struct FirParams {
  const void* in;          // input elements (int8 pairs / float2 / float)
  void* out;               // float (AM/FM or real data) or float2
  const float* taps;       // raw device taps: T floats (real) or T float2 (complex); may be null if tables given
  const float* tapTable;   // rows path: hT[D][tapStride(MP)], pre-scaled; null -> built from `taps` in the prologue
  const float2* mixTable;  // rows path: W[D] (already scaled by inScale); null -> computed from phaseStep
  const float2* rotTable;  // rows path: exp(j*w*m*D), m <= tapStride(MP); null -> computed from phaseStep
  unsigned long long nOut;
  unsigned long long nIn;  // valid input elements starting at `in`
  unsigned long long firstIndex;  // absolute sample index of in[0] (mixer phase)
  unsigned long long phaseStep;   // 2^64 * frac(f/fs)
  unsigned T, D, M;
  unsigned rowsPerTile, outPerTile;
  int mod;        // ModKind of the epilogue
  float gain;     // FM
  float inScale;  // 1/128 for int8 input, 1 otherwise
  // direct kernel only: blockIdx.y selects one of several independent streams (batched audio FIR of the channelizer)
  unsigned long long inBatchStride, outBatchStride;  // in input / output elements
  unsigned winTapChunk;  // window kernel only: taps per phase staged together (multiple of 8; set by the launcher)
  unsigned winStreams;   // window kernel only: number of batched streams (set by the launcher)
};

#ifdef __CUDACC__

__device__ __forceinline__ float2 phasorOfTurns(unsigned long long turns) {
  // turns is a 0.64 fixed-point fraction of a revolution; the signed view keeps |x| <= 1 for sincospif
  const float x = static_cast<float>(static_cast<long long>(turns)) * (2.0f / 18446744073709551616.0f);
  float s, c;
  sincospif(x, &s, &c);
  return make_float2(c, s);
}

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y));
}

__device__ __forceinline__ void storeDemod(int mod, void* out, unsigned long long k, float2 cur, float2 next, float2 rot1, float2 carrier, float gain) {
  if (mod == kModAm) {
    static_cast<float*>(out)[k] = sqrtf(fmaf(cur.x, cur.x, cur.y * cur.y));
  } else if (mod == kModFm) {
    const float2 d = make_float2(fmaf(next.y, cur.y, next.x * cur.x), fmaf(next.y, cur.x, -next.x * cur.y));
    const float2 r = cmulf(d, rot1);
    static_cast<float*>(out)[k] = gain * atan2f(r.y, r.x);
  } else {
    static_cast<float2*>(out)[k] = cmulf(cur, carrier);
  }
}

// Packed FP32 (FFMA2 / FMUL2 / FADD2 on sm_100a): one issue slot carries two FMAs, and the scalar operand is
// broadcast by the instruction itself (FFMA2 Rd, Ra.F32, Rb.F32x2, Rc.F32x2).  The rows kernel is bound by
// instruction issue, not by the FMA pipe, so every per-sample operation works on an (I, Q) register pair.
__device__ __forceinline__ float2 axpy2(float a, float2 x, float2 y) { return __ffma2_rn(make_float2(a, a), x, y); }
__device__ __forceinline__ float2 scale2(float a, float2 x) { return __fmul2_rn(make_float2(a, a), x); }
// z * w with w given as (wx, wy, -wy, wx): z.x*(wx, wy) + z.y*(-wy, wx)
__device__ __forceinline__ float2 cmulPacked(float2 z, float4 w) {
  return axpy2(z.y, make_float2(w.z, w.w), scale2(z.x, make_float2(w.x, w.y)));
}
// Two (I, Q) samples packed as four int8 in one word (already XORed with 0x80808080) -> two exact float pairs
__device__ __forceinline__ void biasedWordToPairs(uint32_t w, float2& z0, float2& z1) {
  constexpr uint32_t kMagic = 0x4B000000u;  // 2^23
  const float2 bias = make_float2(-8388736.0f, -8388736.0f);  // -(2^23 + 128)
  z0 = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(w, kMagic, 0x7650)), __uint_as_float(__byte_perm(w, kMagic, 0x7651))), bias);
  z1 = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(w, kMagic, 0x7652)), __uint_as_float(__byte_perm(w, kMagic, 0x7653))), bias);
}

// Sign-extending byte extraction + I2FP: int8 -> float entirely on the ALU pipe, leaving the FMA pipe (the busiest
// unit of the main loop) to the mixer and the taps.  prmt's selector nibbles with bit 3 set replicate the sign of the
// selected byte; __byte_perm masks that bit off, hence the inline PTX.
template <unsigned SEL>
__device__ __forceinline__ float signedByteToFloat(uint32_t w) {
  int v;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(v) : "r"(w), "n"(SEL));
  return __int2float_rn(v);
}
__device__ __forceinline__ void wordToPairsAlu(uint32_t w, float2& z0, float2& z1) {
  z0 = make_float2(signedByteToFloat<0x8880u>(w), signedByteToFloat<0x9991u>(w));
  z1 = make_float2(signedByteToFloat<0xAAA2u>(w), signedByteToFloat<0xBBB3u>(w));
}

// CONV = 0: magic-number route (PRMT + FADD2, needs the XOR);  CONV = 1: sign-extend + I2FP (ALU pipe only)
template <int CONV>
__device__ __forceinline__ void convertWord(uint32_t w, float2& z0, float2& z1) {
  if constexpr (CONV == 0) {
    biasedWordToPairs(w ^ 0x80808080u, z0, z1);
  } else {
    wordToPairsAlu(w, z0, z1);
  }
}
#ifndef B200SDR_ROWS_CONV
#define B200SDR_ROWS_CONV 1
#endif

// ------------------------------------------------------------------------------------------------
// rows kernel
// ------------------------------------------------------------------------------------------------
template <int ELEM>
struct ElemTraits;
template <>
struct ElemTraits<kElemInt8Complex> {
  static constexpr int kBytes = 2, kVec = 8;
};
template <>
struct ElemTraits<kElemComplex> {
  static constexpr int kBytes = 8, kVec = 2;
};

constexpr int kRowsThreads = 128;

// Row stride (floats) of the transposed tap table hT[p][.] for MP partial sums: next power of two.
__host__ __device__ constexpr int tapStride(int MP) { return MP <= 1 ? 1 : MP <= 2 ? 2 : MP <= 4 ? 4 : MP <= 8 ? 8 : MP <= 16 ? 16 : 32; }

template <int MP>
__device__ __forceinline__ void loadTapRow(const float* hT, unsigned p, float (&h)[MP]) {
  constexpr int TS = tapStride(MP);
  if constexpr (TS == 1) {
    h[0] = hT[p];
  } else if constexpr (TS == 2) {
    const float2 v = reinterpret_cast<const float2*>(hT)[p];
    h[0] = v.x;
    h[1] = v.y;
  } else {
    float t[TS];
#pragma unroll
    for (int i = 0; i < TS / 4; i++) {
      const float4 v = reinterpret_cast<const float4*>(hT)[p * (TS / 4) + i];
      t[4 * i] = v.x;
      t[4 * i + 1] = v.y;
      t[4 * i + 2] = v.z;
      t[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int m = 0; m < MP; m++) h[m] = t[m];
  }
}

// Distance between two rows of the staged tile.  A thread reads ITS row 16 bytes at a time, so rows whose length is a multiple of
// 64 bytes would put a quarter-warp on 2 to 8 of the 8 bank groups; those tiles are staged row by row (one bulk copy per row, issued
// by the thread that owns it) with 16 bytes of padding, which makes the stride odd in 16-byte units.
__host__ __device__ constexpr unsigned rowsRowStride(unsigned D, unsigned elemBytes) {
  return D * elemBytes + ((D * elemBytes >= 128u && (D * elemBytes) % 64u == 0u) ? 16u : 0u);
}

// Shared-memory carve-up (bytes), identical on host and device.
struct RowsSmem {
  unsigned mixOff, rotOff, tapOff, tileOff, total;
};
__host__ __device__ inline RowsSmem rowsSmemLayout(unsigned D, unsigned TS, unsigned M, unsigned rowsPerTile, unsigned elemBytes, bool fm) {
  RowsSmem s;
  unsigned off = 16;  // mbarrier
  s.mixOff = off;
  off += D * 16;  // (wx, wy, -wy, wx) per phase
  s.rotOff = off;
  off += (TS + 1) * 8;
  off = (off + 15u) & ~15u;
  s.tapOff = off;
  off += D * TS * 4;
  off = (off + 127u) & ~127u;
  s.tileOff = off;
  const unsigned tileBytes = rowsPerTile * rowsRowStride(D, elemBytes);
  const unsigned exchBytes = (M > 0 ? (M - 1) : 0) * rowsPerTile * 8 + (fm ? rowsPerTile * 8 : 0);
  off += tileBytes > exchBytes ? tileBytes : exchBytes;
  s.total = off;
  return s;
}

template <int ELEM, bool MIX, int MP, int RPT>
__global__ void __launch_bounds__(kRowsThreads) rowsKernel(const FirParams prm) {
  using Traits = ElemTraits<ELEM>;
  constexpr int ES = Traits::kBytes;
  constexpr int VEC = Traits::kVec;
  extern __shared__ __align__(128) unsigned char smem[];

  const unsigned D = prm.D, M = prm.M, NT = prm.rowsPerTile;
  constexpr int TS = tapStride(MP);
  const RowsSmem lay = rowsSmemLayout(D, TS, M, NT, ES, prm.mod == kModFm);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float4* W = reinterpret_cast<float4*>(smem + lay.mixOff);  // (wx, wy, -wy, wx): both operand pairs of the packed complex multiply
  float2* rot = reinterpret_cast<float2*>(smem + lay.rotOff);
  float* hT = reinterpret_cast<float*>(smem + lay.tapOff);
  unsigned char* tile = smem + lay.tileOff;

  const unsigned tid = threadIdx.x;
  const unsigned long long row0 = static_cast<unsigned long long>(blockIdx.x) * prm.outPerTile;

  // ---- stage the raw tile: one TMA bulk copy, or one per row when the rows are padded -----------------------------
  const unsigned long long tileStart = row0 * D * ES;
  const unsigned long long totalBytes = prm.nIn * ES;
  const unsigned tileBytes = NT * D * ES;
  const unsigned avail = totalBytes - tileStart < tileBytes ? static_cast<unsigned>(totalBytes - tileStart) : tileBytes;
  const unsigned bulk = avail & ~15u;
  const unsigned rowBytes = D * ES, rowStride = rowsRowStride(D, ES);
  const unsigned char* gTile = static_cast<const unsigned char*>(prm.in) + tileStart;
  if (tid == 0) {
    mbarInit(bar, 1);
    fenceMbarInit();
  }
  if (rowStride == rowBytes) {
    if (tid == 0) {
      mbarExpectTx(bar, bulk);
      if (bulk) tmaBulkLoad(tile, gTile, bulk, bar);
    }
    // tail of the last tile: copy the <16 B remainder and zero what lies past the valid input
    for (unsigned b = bulk + tid; b < tileBytes; b += kRowsThreads) tile[b] = b < avail ? gTile[b] : 0;
  } else {
    __syncthreads();  // the barrier is initialised before any thread's copy names it
    if (tid == 0) mbarExpectTx(bar, bulk);  // rowBytes is a multiple of 16: the rows' bulk parts add up to `bulk`
#pragma unroll
    for (int i = 0; i < RPT; i++) {
      const unsigned row = tid + i * kRowsThreads, start = row * rowBytes;
      const unsigned have = avail > start ? (avail - start < rowBytes ? avail - start : rowBytes) : 0u;
      const unsigned rb = have & ~15u;
      if (rb) tmaBulkLoad(tile + row * rowStride, gTile + start, rb, bar);
      for (unsigned b = rb; b < rowBytes; b++) tile[row * rowStride + b] = b < have ? gTile[start + b] : 0;
    }
  }

  // ---- tables ----------------------------------------------------------------------------------------
  for (unsigned i = tid; i < D * TS; i += kRowsThreads) {
    float v;
    if (prm.tapTable) {
      v = prm.tapTable[i];
    } else {
      const unsigned p = i / TS, m = i % TS;
      const unsigned j = m * D + p;
      v = (m < M && j < prm.T) ? prm.taps[j] * (MIX ? 1.0f : prm.inScale) : 0.0f;
    }
    hT[i] = v;
  }
  if constexpr (MIX) {
    for (unsigned p = tid; p < D; p += kRowsThreads) {
      float2 w;
      if (prm.mixTable) {
        w = prm.mixTable[p];
      } else {
        w = phasorOfTurns(prm.phaseStep * p);
        w.x *= prm.inScale;
        w.y *= prm.inScale;
      }
      W[p] = make_float4(w.x, w.y, -w.y, w.x);
    }
    if (tid <= TS) {
      rot[tid] = prm.rotTable ? prm.rotTable[tid] : phasorOfTurns(prm.phaseStep * (static_cast<unsigned long long>(tid) * D));
    }
  }
  __syncthreads();
  mbarWait(bar, 0);

  // ---- main loop: convert + mix once per sample, M partial sums per row ----------------------------------
  float2 acc[RPT][MP];
#pragma unroll
  for (int i = 0; i < RPT; i++)
#pragma unroll
    for (int m = 0; m < MP; m++) acc[i][m] = make_float2(0.0f, 0.0f);

  const unsigned char* rowPtr[RPT];
#pragma unroll
  for (int i = 0; i < RPT; i++) rowPtr[i] = tile + (tid + i * kRowsThreads) * rowStride;

  // wide variants: two iterations in flight, so that the next sample pair's tap rows (MP / 2 LDS.128) load under this pair's FMAs
#pragma unroll(MP >= 16 ? 2 : 1)
  for (unsigned p = 0; p < D; p += VEC) {
    uint4 v[RPT];
#pragma unroll
    for (int i = 0; i < RPT; i++) v[i] = *reinterpret_cast<const uint4*>(rowPtr[i] + p * ES);

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
#pragma unroll
        for (int i = 0; i < RPT; i++) {
          const uint32_t word = s == 0 ? v[i].x : s == 1 ? v[i].y : s == 2 ? v[i].z : v[i].w;
          float2 z0, z1;
          convertWord<B200SDR_ROWS_CONV>(word, z0, z1);
          if constexpr (MIX) {
            z0 = cmulPacked(z0, w0);
            z1 = cmulPacked(z1, w1);
          }
#pragma unroll
          for (int m = 0; m < MP; m++) {
            acc[i][m] = axpy2(h0[m], z0, acc[i][m]);
            acc[i][m] = axpy2(h1[m], z1, acc[i][m]);
          }
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
#pragma unroll
      for (int i = 0; i < RPT; i++) {
        float2 z0 = make_float2(__uint_as_float(v[i].x), __uint_as_float(v[i].y));
        float2 z1 = make_float2(__uint_as_float(v[i].z), __uint_as_float(v[i].w));
        if constexpr (MIX) {
          z0 = cmulPacked(z0, w0);
          z1 = cmulPacked(z1, w1);
        }
#pragma unroll
        for (int m = 0; m < MP; m++) {
          acc[i][m] = axpy2(h0[m], z0, acc[i][m]);
          acc[i][m] = axpy2(h1[m], z1, acc[i][m]);
        }
      }
    }
  }

  // ---- exchange partial sums: y[k] = P[k][0] + sum_{m>=1} rot[m] * P[k+m][m] -------------------------
  float2* part = reinterpret_cast<float2*>(tile);       // [(m-1)*NT + row]
  float2* sums = part + (M > 0 ? (M - 1) : 0) * NT;     // FM only: S[row]
  if (MP > 1 && M > 1) {
    __syncthreads();  // everyone is done reading the tile; it is reused for the exchange
#pragma unroll
    for (int i = 0; i < RPT; i++) {
      const unsigned row = tid + i * kRowsThreads;
#pragma unroll
      for (int m = 1; m < MP; m++) {
        if (m < M) {
          float2 v = acc[i][m];
          if constexpr (MIX) v = cmulf(v, rot[m]);
          part[(m - 1) * NT + row] = v;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RPT; i++) {
      const unsigned row = tid + i * kRowsThreads;
#pragma unroll
      for (int m = 1; m < MP; m++) {
        if (m < M && row + m < NT) {
          const float2 v = part[(m - 1) * NT + row + m];
          acc[i][0].x += v.x;
          acc[i][0].y += v.y;
        }
      }
    }
  }

  float2 rot1 = make_float2(1.0f, 0.0f);
  if constexpr (MIX) rot1 = rot[1];

  if (prm.mod == kModFm) {
    if (!(MP > 1 && M > 1)) __syncthreads();
#pragma unroll
    for (int i = 0; i < RPT; i++) sums[tid + i * kRowsThreads] = acc[i][0];
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < RPT; i++) {
    const unsigned row = tid + i * kRowsThreads;
    const unsigned long long k = row0 + row;
    if (row < prm.outPerTile && k < prm.nOut) {
      float2 next = make_float2(0.0f, 0.0f), carrier = make_float2(1.0f, 0.0f);
      if (prm.mod == kModFm) next = sums[row + 1];
      if (MIX && prm.mod == kModNone) carrier = phasorOfTurns(prm.phaseStep * (prm.firstIndex + k * D));
      storeDemod(prm.mod, prm.out, k, acc[i][0], next, rot1, carrier, prm.gain);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// direct kernel (fallback): thread per output, taps chunked through shared memory as float2 (real
// taps carry a zero imaginary part unless the mixer phasor is folded in).  When the block's input
// window ((outPerBlock-1)*D + T elements) fits in shared memory it is staged there with coalesced loads
// (STAGED), which is what the audio-rate FIR of the chain uses; otherwise inputs come through L1.
// ------------------------------------------------------------------------------------------------
constexpr int kDirectThreads = 256;
constexpr int kDirectTapChunk = 1024;
constexpr unsigned kDirectFixedSmem = kDirectTapChunk * 8 + kDirectThreads * 8;

template <int ELEM>
struct RawElem;
template <>
struct RawElem<kElemInt8Complex> {
  using type = char2;
};
template <>
struct RawElem<kElemComplex> {
  using type = float2;
};
template <>
struct RawElem<kElemReal> {
  using type = float;
};

template <int ELEM>
__device__ __forceinline__ float2 elemToComplex(typename RawElem<ELEM>::type v, float scale) {
  if constexpr (ELEM == kElemInt8Complex) {
    return make_float2(static_cast<float>(v.x) * scale, static_cast<float>(v.y) * scale);
  } else if constexpr (ELEM == kElemComplex) {
    return v;
  } else {
    return make_float2(v, 0.0f);
  }
}

// TAPC: taps are complex (FirCC/FirCF).  REALOUT: real data with real taps -> float output.
template <int ELEM, bool TAPC, bool MIX, bool REALOUT, bool STAGED>
__global__ void __launch_bounds__(kDirectThreads) directKernel(const FirParams prm) {
  using Raw = typename RawElem<ELEM>::type;
  extern __shared__ __align__(16) unsigned char dsmem[];
  float2* sTaps = reinterpret_cast<float2*>(dsmem);
  float2* sY = sTaps + kDirectTapChunk;
  Raw* sIn = reinterpret_cast<Raw*>(dsmem + kDirectFixedSmem);

  const unsigned tid = threadIdx.x;
  const unsigned outPerBlock = prm.mod == kModFm ? kDirectThreads - 1 : kDirectThreads;
  const unsigned long long k0 = static_cast<unsigned long long>(blockIdx.x) * outPerBlock;
  const unsigned long long k = k0 + tid;
  // FM needs FIR output k+1 as well: the block computes one more FIR output than it stores.
  const unsigned long long nFir = prm.mod == kModFm ? prm.nOut + 1 : prm.nOut;
  const bool active = k < nFir;
  const unsigned long long base = k * prm.D;
  const Raw* gIn = static_cast<const Raw*>(prm.in) + blockIdx.y * prm.inBatchStride;

  if constexpr (STAGED) {
    const unsigned long long first = k0 * prm.D;
    const unsigned tileElems = (kDirectThreads - 1) * prm.D + prm.T;
    for (unsigned i = tid; i < tileElems; i += kDirectThreads) {
      Raw v {};
      if (first + i < prm.nIn) v = gIn[first + i];
      sIn[i] = v;
    }
  }

  float2 acc = make_float2(0.0f, 0.0f);
  for (unsigned c0 = 0; c0 < prm.T; c0 += kDirectTapChunk) {
    const unsigned cn = prm.T - c0 < kDirectTapChunk ? prm.T - c0 : kDirectTapChunk;
    __syncthreads();
    for (unsigned j = tid; j < cn; j += kDirectThreads) {
      float2 h;
      if constexpr (TAPC) {
        h = reinterpret_cast<const float2*>(prm.taps)[c0 + j];
      } else {
        h = make_float2(prm.taps[c0 + j], 0.0f);
      }
      if constexpr (MIX) h = cmulf(h, phasorOfTurns(prm.phaseStep * (c0 + j)));
      sTaps[j] = h;
    }
    __syncthreads();
    if (active) {
      const Raw* src = STAGED ? sIn + (tid * prm.D + c0) : gIn + (base + c0);
#pragma unroll 4
      for (unsigned j = 0; j < cn; j++) {
        const float2 x = elemToComplex<ELEM>(src[j], prm.inScale);
        const float2 h = sTaps[j];
        if constexpr (ELEM == kElemReal) {
          acc.x = fmaf(h.x, x.x, acc.x);
          if constexpr (TAPC) acc.y = fmaf(h.y, x.x, acc.y);
        } else if constexpr (TAPC || MIX) {
          acc.x = fmaf(h.x, x.x, acc.x);
          acc.x = fmaf(-h.y, x.y, acc.x);
          acc.y = fmaf(h.x, x.y, acc.y);
          acc.y = fmaf(h.y, x.x, acc.y);
        } else {
          acc.x = fmaf(h.x, x.x, acc.x);
          acc.y = fmaf(h.x, x.y, acc.y);
        }
      }
    }
  }

  if constexpr (REALOUT) {
    if (active) static_cast<float*>(prm.out)[blockIdx.y * prm.outBatchStride + k] = acc.x;
    return;
  }

  float2 rot1 = make_float2(1.0f, 0.0f), carrier = make_float2(1.0f, 0.0f), next = make_float2(0.0f, 0.0f);
  if constexpr (MIX) rot1 = phasorOfTurns(prm.phaseStep * static_cast<unsigned long long>(prm.D));
  if (prm.mod == kModFm) {
    sY[tid] = acc;
    __syncthreads();
    if (tid + 1 < kDirectThreads) next = sY[tid + 1];
  }
  if (tid < outPerBlock && k < prm.nOut) {
    if (MIX && prm.mod == kModNone) carrier = phasorOfTurns(prm.phaseStep * (prm.firstIndex + base));
    void* out = prm.mod == kModNone ? static_cast<void*>(static_cast<float2*>(prm.out) + blockIdx.y * prm.outBatchStride)
                                    : static_cast<void*>(static_cast<float*>(prm.out) + blockIdx.y * prm.outBatchStride);
    storeDemod(prm.mod, out, k, acc, next, rot1, carrier, prm.gain);
  }
}

#endif  // __CUDACC__

}  // namespace b200sdr
