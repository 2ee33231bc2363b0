// chainKernel -- the whole int8/cf32 -> [mix] -> decimating FIR -> AM/FM demod -> audio FIR chain in ONE persistent
// kernel (sm_100a).  HBM traffic is the algorithmic minimum: every input byte is read once (plus a small halo per
// CTA), only the final audio samples are written; the demodulated stream never leaves shared memory.
//
//   * Each CTA owns a contiguous run of audio outputs and walks the input rows they need tile by tile.  Raw tiles
//     (rows x D1 samples) are staged in shared memory by 1-D TMA bulk copies through a ring of `stages` buffers, each
//     with its own mbarrier, so the copy of tile t+stages overlaps the arithmetic of tiles t .. t+stages-1.
//   * The per-tile arithmetic is that of rowsKernel (fir_kernels.cuh): convert + mix once per sample, M partial sums
//     per row in packed FP32, one exchange of the partial sums through shared memory.
//   * The demodulated samples of a tile are appended to a small shared-memory line; every audio output whose T2-sample
//     window is complete is computed from it (threads split each dot product `audioParts` ways and combine with
//     shuffles), then the unconsumed tail is carried over to the next tile.
//   * Taps, mixer phasors and rotations are loaded once per CTA.
#pragma once

#include "fir_kernels.cuh"

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
  unsigned rowsPerTile, outPerTile;
  unsigned stages;         // TMA ring depth
  unsigned audioParts;     // threads per audio dot product (power of two <= 32)
  unsigned dmCapacity;     // floats per demod line
  int mod;                 // kModAm / kModFm
  float gain;
};

struct ChainSmem {
  unsigned mixOff, rotOff, tapOff, taps2Off, partOff, dmOff, tileOff, tileBytes, total;
};

__host__ __device__ inline ChainSmem chainSmemLayout(unsigned D, unsigned TS, unsigned M, unsigned rowsPerTile, unsigned elemBytes,
                                                     bool fm, unsigned T2, unsigned dmCapacity, unsigned stages) {
  ChainSmem s;
  unsigned off = 64;  // up to 8 mbarriers
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
  s.partOff = off;
  off += ((M > 0 ? M - 1 : 0) + (fm ? 1u : 0u)) * rowsPerTile * 8;
  off = (off + 15u) & ~15u;
  s.dmOff = off;
  off += 2u * dmCapacity * 4;
  off = (off + 127u) & ~127u;
  s.tileOff = off;
  s.tileBytes = rowsPerTile * D * elemBytes;
  off += stages * s.tileBytes;
  s.total = off;
  return s;
}

#ifdef __CUDACC__

__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One tile: RPT rows per thread, MP partial sums per row.  `tile` is the staged raw tile in shared memory.
template <int ELEM, bool MIX, int MP, int RPT, int CONV>
__device__ __forceinline__ void tilePartialSums(const unsigned char* tile, const float* hT, const float4* W, unsigned D, unsigned tid,
                                                float2 (&acc)[RPT][MP]) {
  constexpr int ES = ElemTraits<ELEM>::kBytes;
  constexpr int VEC = ElemTraits<ELEM>::kVec;
#pragma unroll
  for (int i = 0; i < RPT; i++)
#pragma unroll
    for (int m = 0; m < MP; m++) acc[i][m] = make_float2(0.0f, 0.0f);
  const unsigned rowBytes = D * ES;
  const unsigned char* rowPtr[RPT];
#pragma unroll
  for (int i = 0; i < RPT; i++) rowPtr[i] = tile + (tid + i * kRowsThreads) * rowBytes;

  for (unsigned p = 0; p < D; p += VEC) {
    uint4 v[RPT];
#pragma unroll
    for (int i = 0; i < RPT; i++) v[i] = *reinterpret_cast<const uint4*>(rowPtr[i] + p * ES);
    if constexpr (ELEM == kElemInt8Complex) {
#pragma unroll
      for (int s = 0; s < 4; s++) {
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
          convertWord<CONV>(word, z0, z1);
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
}

template <int ELEM, bool MIX, int MP, int RPT, int CONV>
__global__ void __launch_bounds__(kRowsThreads) chainKernel(const ChainParams prm) {
  constexpr int ES = ElemTraits<ELEM>::kBytes;
  constexpr int TS = tapStride(MP);
  extern __shared__ __align__(128) unsigned char smem[];

  const unsigned D = prm.D1, M = prm.M, NT = prm.rowsPerTile, OT = prm.outPerTile, S = prm.stages;
  const unsigned T2 = prm.T2, D2 = prm.D2;
  const bool fm = prm.mod == kModFm;
  const ChainSmem lay = chainSmemLayout(D, TS, M, NT, ES, fm, T2, prm.dmCapacity, S);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  float4* W = reinterpret_cast<float4*>(smem + lay.mixOff);
  float2* rot = reinterpret_cast<float2*>(smem + lay.rotOff);
  float* hT = reinterpret_cast<float*>(smem + lay.tapOff);
  float* h2 = reinterpret_cast<float*>(smem + lay.taps2Off);
  float2* part = reinterpret_cast<float2*>(smem + lay.partOff);       // [(m-1)*NT + row]
  float2* sums = part + (M > 0 ? (M - 1) : 0) * NT;                    // FM only
  float* dm = reinterpret_cast<float*>(smem + lay.dmOff);             // two lines of dmCapacity floats
  unsigned char* tiles = smem + lay.tileOff;

  const unsigned tid = threadIdx.x;

  // ---- this CTA's run of audio outputs ---------------------------------------------------------------
  const unsigned long long per = prm.nAudio / gridDim.x, extra = prm.nAudio % gridDim.x;
  const unsigned long long a0 = blockIdx.x * per + (blockIdx.x < extra ? blockIdx.x : extra);
  const unsigned long long cnt = per + (blockIdx.x < extra ? 1 : 0);
  if (cnt == 0) return;
  const unsigned long long row0 = a0 * D2;                       // first RF row (= demod index) this CTA needs
  const unsigned long long needDemod = (cnt - 1) * D2 + T2;
  const unsigned nTiles = static_cast<unsigned>((needDemod + OT - 1) / OT);
  const unsigned long long totalBytes = prm.nIn * ES;
  const unsigned tileBytes = lay.tileBytes;
  const unsigned char* gin = static_cast<const unsigned char*>(prm.in);

  auto tileAvail = [&](unsigned t) -> unsigned {
    const unsigned long long start = (row0 + static_cast<unsigned long long>(t) * OT) * D * ES;
    if (start >= totalBytes) return 0u;
    const unsigned long long left = totalBytes - start;
    return left < tileBytes ? static_cast<unsigned>(left) : tileBytes;
  };
  // issue the bulk copy of tile t into its ring slot (one thread)
  auto issueTile = [&](unsigned t) {
    const unsigned slot = t % S;
    const unsigned bulk = tileAvail(t) & ~15u;
    mbarExpectTx(&bars[slot], bulk);
    if (bulk) tmaBulkLoad(tiles + slot * tileBytes, gin + (row0 + static_cast<unsigned long long>(t) * OT) * D * ES, bulk, &bars[slot]);
  };
  // bytes of tile t that the bulk copy does not cover (the <16 B remainder and everything past the valid input)
  auto fillTileTail = [&](unsigned t) {
    const unsigned avail = tileAvail(t);
    if (avail == tileBytes) return;
    const unsigned bulk = avail & ~15u;
    unsigned char* dst = tiles + (t % S) * tileBytes;
    const unsigned long long start = (row0 + static_cast<unsigned long long>(t) * OT) * D * ES;
    for (unsigned b = bulk + tid; b < tileBytes; b += kRowsThreads) dst[b] = b < avail ? gin[start + b] : 0;
  };

  if (tid == 0) {
    for (unsigned s = 0; s < S; s++) mbarInit(&bars[s], 1);
    fenceMbarInit();
    for (unsigned t = 0; t < S && t < nTiles; t++) issueTile(t);
  }
  for (unsigned t = 0; t < S && t < nTiles; t++) fillTileTail(t);

  // ---- tables (once per CTA) ---------------------------------------------------------------------------
  for (unsigned i = tid; i < D * TS; i += kRowsThreads) hT[i] = prm.tapTable[i];
  for (unsigned i = tid; i < ((T2 + 3u) & ~3u); i += kRowsThreads) h2[i] = i < T2 ? prm.taps2[i] : 0.0f;
  if constexpr (MIX) {
    for (unsigned p = tid; p < D; p += kRowsThreads) {
      const float2 w = prm.mixTable[p];
      W[p] = make_float4(w.x, w.y, -w.y, w.x);
    }
    if (tid <= TS) rot[tid] = prm.rotTable[tid];
  }
  __syncthreads();

  float2 rot1 = make_float2(1.0f, 0.0f);
  if constexpr (MIX) rot1 = rot[1];

  unsigned carry = 0, cur = 0;
  unsigned long long done = 0;
  const unsigned P = prm.audioParts;
  const unsigned groups = kRowsThreads / P;
  const unsigned chunk = ((T2 + P - 1) / P + 3u) & ~3u;  // taps per part, a multiple of 4 (aligned 128-bit tap loads)
  const bool pairs = (D2 & 1u) == 0;

  for (unsigned t = 0; t < nTiles; t++) {
    const unsigned slot = t % S;
    mbarWait(&bars[slot], (t / S) & 1u);

    float2 acc[RPT][MP];
    tilePartialSums<ELEM, MIX, MP, RPT, CONV>(tiles + slot * tileBytes, hT, W, D, tid, acc);

    // ---- exchange partial sums: y[k] = P[k][0] + sum_{m>=1} rot[m] * P[k+m][m] ------------------------------
    if (MP > 1 && M > 1) {
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
    }
    __syncthreads();  // B1: every thread is done with the tile slot; partial sums are visible

    if (t + S < nTiles) {  // refill the slot while the rest of this tile and the next tiles are processed
      if (tid == 0) {
        fenceProxyAsync();
        issueTile(t + S);
      }
      fillTileTail(t + S);
    }

    if (MP > 1 && M > 1) {
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

    float* line = dm + cur * prm.dmCapacity;
    if (fm) {
#pragma unroll
      for (int i = 0; i < RPT; i++) sums[tid + i * kRowsThreads] = acc[i][0];
      __syncthreads();
#pragma unroll
      for (int i = 0; i < RPT; i++) {
        const unsigned row = tid + i * kRowsThreads;
        if (row < OT) {
          const float2 c = acc[i][0], n = sums[row + 1];
          const float2 d = make_float2(fmaf(n.y, c.y, n.x * c.x), fmaf(n.y, c.x, -n.x * c.y));
          const float2 r = cmulf(d, rot1);
          line[carry + row] = prm.gain * atan2f(r.y, r.x);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < RPT; i++) {
        const unsigned row = tid + i * kRowsThreads;
        if (row < OT) line[carry + row] = sqrtf(fmaf(acc[i][0].x, acc[i][0].x, acc[i][0].y * acc[i][0].y));
      }
    }
    __syncthreads();  // B2: the demod line is complete

    // ---- audio FIR over the line; carry the unconsumed tail into the other line ------------------------------
    const unsigned len = carry + OT;
    unsigned nA = len >= T2 ? (len - T2) / D2 + 1 : 0;
    if (static_cast<unsigned long long>(nA) > cnt - done) nA = static_cast<unsigned>(cnt - done);
    const unsigned group = tid / P, partIdx = tid % P;
    const unsigned j0 = partIdx * chunk, j1 = j0 + chunk < T2 ? j0 + chunk : T2;
    for (unsigned base = 0; base < nA; base += groups) {
      const unsigned o = base + group;
      float y = 0.0f;
      if (o < nA && j0 < T2) {
        const float* x = line + o * D2 + j0;
        const float* h = h2 + j0;  // 16-byte aligned: chunk is a multiple of 4
        const unsigned n = j1 - j0;
        float y0 = 0.0f, y1 = 0.0f, y2 = 0.0f, y3 = 0.0f;
        unsigned j = 0;
        if (pairs) {  // o*D2 + j0 is even: 64-bit loads of the demod line
#pragma unroll 4
          for (; j + 4 <= n; j += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(h + j);
            const float2 xa = *reinterpret_cast<const float2*>(x + j);
            const float2 xb = *reinterpret_cast<const float2*>(x + j + 2);
            y0 = fmaf(hv.x, xa.x, y0);
            y1 = fmaf(hv.y, xa.y, y1);
            y2 = fmaf(hv.z, xb.x, y2);
            y3 = fmaf(hv.w, xb.y, y3);
          }
        } else {
#pragma unroll 4
          for (; j + 4 <= n; j += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(h + j);
            y0 = fmaf(hv.x, x[j], y0);
            y1 = fmaf(hv.y, x[j + 1], y1);
            y2 = fmaf(hv.z, x[j + 2], y2);
            y3 = fmaf(hv.w, x[j + 3], y3);
          }
        }
        for (; j < n; j++) y0 = fmaf(h[j], x[j], y0);
        y = (y0 + y1) + (y2 + y3);
      }
      for (unsigned off = P >> 1; off; off >>= 1) y += __shfl_xor_sync(0xffffffffu, y, off);
      if (partIdx == 0 && o < nA) prm.out[a0 + done + o] = y;
    }
    const unsigned consumed = nA * D2;
    const unsigned newCarry = len > consumed ? len - consumed : 0;
    float* next = dm + (cur ^ 1u) * prm.dmCapacity;
    for (unsigned i = tid; i < newCarry; i += kRowsThreads) next[i] = line[consumed + i];
    done += nA;
    carry = newCarry;
    cur ^= 1u;
  }
}

#endif  // __CUDACC__

}  // namespace b200sdr
