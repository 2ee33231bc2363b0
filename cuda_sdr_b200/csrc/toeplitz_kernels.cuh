// toepKernel -- the int8 -> mix -> decimating FIR -> AM/FM demod -> audio FIR chain as ONE persistent kernel whose RF
// stage is a dense int8 contraction over a Toeplitz view of the raw input (sm_100a).
//
// The decimating FIR with the mixer folded into complex taps c[j] = h[j] * exp(j*w*j) / 128 is
//     y[k] = sum_j c[j] * z[k*D + j],   z = I + jQ  (the per-output carrier exp(j*w*k*D) drops out of |y| and of the
//                                                    FM discriminator up to the constant rotation rot1 = exp(j*w*D)).
// Take the raw byte stream x (I,Q,I,Q,...) and view it as a matrix A whose row i starts at byte i * 8D (four outputs
// per A-row) and is K = 2*T + 6*D bytes long -- rows overlap, nothing is copied.  With
//     B[2*(D*n + j) + 0][2n + 0] =  re c[j]    B[2*(D*n + j) + 0][2n + 1] = im c[j]
//     B[2*(D*n + j) + 1][2n + 0] = -im c[j]    B[2*(D*n + j) + 1][2n + 1] = re c[j]        (n = 0..3, j = 0..T-1)
// the product (A B)[i][2n + e] is re/im of y[4i + n]: ONE int8 x int8 -> int32 GEMM per tile does convert + mix + FIR +
// decimate, with no partial sums to exchange and no halo between warps beyond the taps' own overlap.  B is held in 24-bit
// fixed point as three signed int8 digits (three exact IMMA.16832.S8.S8 per k-step; error <= 2^-24 of the largest entry).
//
// Fragment mapping of mma.sync.m16n8k32 (g = lane / 4, t = lane % 4): A-rows are 8D bytes apart, which is 0 or 64 mod 128 --
// in a dense tile every fragment load would hit the same banks.  The tile is therefore staged by a TMA *tensor* copy with
// hardware swizzle: the input is described as a 3-D tensor {W bytes, chunk (stride W), shift (stride 16 B)} (W = 64 or
// 128), so ONE cp.async.bulk.tensor per warp block fetches any 16-byte-aligned run of the stream and lands it densely but
// XOR-swizzled (address bits [4,6) ^= bits [7,9) for W = 64; [4,7) ^= [7,10) for W = 128).  With that swizzle the eight
// row addresses of an ldmatrix (320 B apart for D = 40) fall into eight different 16-byte bank groups, and one
// conflict-free ldmatrix.x4 per (m-tile, k-step) delivers {a0, a1, a2, a3} directly in IMMA register order.
// The accumulator registers of lane L = 4g + t are exactly outputs L (rows g) and 32 + L (rows g + 8) of the 64-output
// m-tile: demodulation and the store into the demod line are perfectly coalesced with no shuffle (AM) or one (FM).
//
// Division of labour: every warp owns a private TMA ring (S slots of one warp block = G m-tiles = 64*G outputs); it refills
// a slot itself as soon as its k-loop has consumed it, so no warp ever waits for another one to issue a copy.  Demodulated
// samples go into a shared-memory ring of 4 tiles (+ a mirror of the first T2 samples past its end, so every FIR window is
// contiguous; no carry copies).  NA audio warps take the tiles in turn (tile t -> audio warp t mod NA): each waits for
// "its" tile to be complete (mbarrier, one arrival per compute warp), runs the audio FIR over the outputs whose window that
// tile completes, and releases the tiles no later window needs (mbarrier).  Consecutive tiles are processed concurrently,
// so the audio stage is no longer the serial stage it is in chainKernel.
#pragma once

#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched from the driver at run time)

#include "chain_kernels.cuh"

namespace b200sdr {

struct ToepParams {
  const unsigned char* in;  // interleaved int8 I,Q; 16-byte aligned
  float* out;               // audio outputs
  const float* taps2;       // T2 audio taps, zero-padded to a multiple of 4
  const uint4* bFrag;       // [Q][lane 32][3] : words (ksub*3 + digit)*2 + half of k-steps 2q, 2q+1 (natural k order)
  unsigned long long nInBytes;
  unsigned long long nAudio;
  unsigned D1, T2, D2;
  unsigned Q, KS;           // pairs of k-steps; k-steps (KS = 2Q or 2Q-1)
  unsigned NW, NA, S;       // compute warps, audio warps, ring depth per compute warp
  unsigned slotBytes;       // ring slot stride (multiple of 1024: swizzle atoms)
  unsigned blockBytes;      // input bytes one warp block reads: (16G-1)*8D + 64Q
  unsigned boxBytes;        // bytes one tensor copy delivers: blockBytes rounded up to W
  unsigned wShift;          // log2 W (6 or 7): inner extent of the tensor map = swizzle span
  unsigned swzMask;         // 0x30 (W = 64) or 0x70 (W = 128): offset ^= (offset >> 3) & swzMask
  unsigned long long tmaEnd;  // input bytes below this are reachable by the tensor map whatever the shift
  unsigned prefetch;        // tiles ahead of the ring's newest slot whose block is prefetched into L2 (0 = off)
  unsigned span;            // tiles before tile t that the audio windows completed by tile t reach into: ceil((T2-1)/OT) <= 2
  int fm;
  float gain;
  float2 rot1;              // exp(j*w*D1) (FM only)
  float s0, s1, s2;         // value = acc0*s0 + acc1*s1 + acc2*s2
};

constexpr unsigned kToepLines = 4;  // tiles in the demod ring

struct ToepSmem {
  unsigned bFragOff, taps2Off, dmOff, slotOff, total;
};

// barriers: tileDone[4] at 0, tileFull[4] at 32, constants at 64, full[NW*S] at 128 (NW*S <= 48)
__host__ __device__ inline ToepSmem toepSmemLayout(unsigned Q, unsigned T2, unsigned OT, unsigned NW, unsigned S, unsigned slotBytes) {
  ToepSmem s;
  unsigned off = 512;
  s.bFragOff = off;
  off += Q * 1536u;
  s.taps2Off = off;
  off += ((T2 + 3u) & ~3u) * 4u;
  s.dmOff = off;
  off += (kToepLines * OT + ((T2 + 3u) & ~3u)) * 4u;  // ring + mirror
  off = (off + 1023u) & ~1023u;
  s.slotOff = off;
  off += NW * S * slotBytes;
  s.total = off;
  return s;
}

#ifdef __CUDACC__

constexpr unsigned kToepMagicBits = 0x4B400000u;  // 1.5 * 2^23: int32 accumulators start here, so their bits ARE the float
constexpr float kToepMagic = 12582912.0f;

template <int G>
struct ToepFrag {
  unsigned a[G][2][4];  // [m-tile][k-step of the pair][a0..a3]
  uint4 b[3];           // 12 words: (ksub*3 + digit)*2 + half
};

__device__ __forceinline__ void ldsm4(unsigned (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <int G>
__device__ __forceinline__ void toepMma(int (&acc)[G][3][4], const ToepFrag<G>& f, bool both) {
  const unsigned w[12] = {f.b[0].x, f.b[0].y, f.b[0].z, f.b[0].w, f.b[1].x, f.b[1].y, f.b[1].z, f.b[1].w, f.b[2].x, f.b[2].y, f.b[2].z, f.b[2].w};
#pragma unroll
  for (int j = 0; j < G; j++)
#pragma unroll
    for (int d = 0; d < 3; d++) imma16832(acc[j][d], f.a[j][0][0], f.a[j][0][1], f.a[j][0][2], f.a[j][0][3], w[2 * d], w[2 * d + 1]);
  if (both) {
#pragma unroll
    for (int j = 0; j < G; j++)
#pragma unroll
      for (int d = 0; d < 3; d++) imma16832(acc[j][d], f.a[j][1][0], f.a[j][1][1], f.a[j][1][2], f.a[j][1][3], w[6 + 2 * d], w[6 + 2 * d + 1]);
  }
}

template <int G, bool MAGIC>
__global__ void __launch_bounds__(384, 1) toepKernel(const ToepParams prm, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const unsigned NW = prm.NW, NA = prm.NA, S = prm.S;
  const unsigned D = prm.D1, T2 = prm.T2, D2 = prm.D2;
  const unsigned fm = prm.fm ? 1u : 0u;
  const unsigned OTW = 64u * G - fm;  // demod outputs per warp block
  const unsigned OT = NW * OTW;       // demod outputs per tile
  const unsigned AS = 8u * D;         // A-row stride (4 outputs)
  const unsigned R = kToepLines * OT; // demod ring, floats
  const unsigned mirror = (T2 + 3u) & ~3u;
  const ToepSmem lay = toepSmemLayout(prm.Q, T2, OT, NW, S, prm.slotBytes);
  uint64_t* tileDone = reinterpret_cast<uint64_t*>(smem);
  uint64_t* constBar = reinterpret_cast<uint64_t*>(smem + 64);
  uint64_t* tileFull = reinterpret_cast<uint64_t*>(smem + 32);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 128);
  const uint4* bFrag = reinterpret_cast<const uint4*>(smem + lay.bFragOff);
  const float* h2 = reinterpret_cast<const float*>(smem + lay.taps2Off);
  float* ring = reinterpret_cast<float*>(smem + lay.dmOff);
  unsigned char* slots = smem + lay.slotOff;

  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;

  // ---- this CTA's run of audio outputs ---------------------------------------------------------------
  const unsigned long long per = prm.nAudio / gridDim.x, extra = prm.nAudio % gridDim.x;
  const unsigned long long a0 = blockIdx.x * per + (blockIdx.x < extra ? blockIdx.x : extra);
  const unsigned long long cnt = per + (blockIdx.x < extra ? 1 : 0);
  if (cnt == 0) return;
  const unsigned long long row0 = a0 * D2;
  const unsigned long long needDemod = (cnt - 1) * D2 + T2;
  const unsigned nTiles = static_cast<unsigned>((needDemod + OT - 1) / OT);
  const unsigned long long totalBytes = prm.nInBytes;
  const unsigned char* gin = prm.in;
  const unsigned blockBytes = prm.blockBytes, slotBytes = prm.slotBytes;

  // ---- prologue ------------------------------------------------------------------------------------------
  if (tid == 0) {
    for (unsigned i = 0; i < NW * S; i++) mbarInit(&full[i], 1);
    for (unsigned i = 0; i < kToepLines; i++) {
      mbarInit(&tileDone[i], prm.span + 1u);
      mbarInit(&tileFull[i], NW);
    }
    mbarInit(constBar, 1);
    fenceMbarInit();
    // constants (written when the chain was created): B fragments and audio taps, one bulk copy each
    mbarExpectTx(constBar, prm.Q * 1536u + mirror * 4u);
    tmaBulkLoad(smem + lay.bFragOff, prm.bFrag, prm.Q * 1536u, constBar);
    tmaBulkLoad(smem + lay.taps2Off, prm.taps2, mirror * 4u, constBar);
  }
  __syncthreads();  // the only CTA-wide barrier
  // Programmatic dependent launch: everything above touches only this CTA's shared memory and constants written when the
  // chain was created, so it may overlap the tail of the previous kernel in the stream; the input may be that kernel's
  // output and the audio buffer may be its input, so every thread waits here before touching either.  (No-ops when the
  // kernel is launched without the attribute.)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // byte offset of warp w's block of tile t
  auto blockStart = [&](unsigned t, unsigned w) { return (row0 + static_cast<unsigned long long>(t) * OT + static_cast<unsigned long long>(w) * OTW) * 2ull * D; };
  // (whole warp) stage the block of tile t into `slot`: one tensor copy, chunk coordinate start / W, shift (start % W) / 16
  auto issueBlock = [&](unsigned t, unsigned slot) {
    __syncwarp();
    if (lane == 0) {
      const unsigned long long start = blockStart(t, warp);
      fenceProxyAsync();  // this warp's generic-proxy reads of the slot come before the async-proxy write
      mbarExpectTx(&full[warp * S + slot], prm.boxBytes);
      const int c1 = static_cast<int>(start >> prm.wShift), c2 = static_cast<int>((start & ((1u << prm.wShift) - 1u)) >> 4);
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                       smemAddr(slots + (warp * S + slot) * slotBytes)),
                   "l"(&tmap), "r"(0), "r"(c1), "r"(c2), "r"(smemAddr(&full[warp * S + slot]))
                   : "memory");
    }
  };
  // (whole warp, after the copy has landed) the last bytes of the input lie in a chunk the tensor map cannot reach for
  // every shift: the copy zero-filled them, put them in by hand
  auto patchBlock = [&](unsigned t, unsigned slot) {
    const unsigned long long start = blockStart(t, warp);
    if (start + blockBytes <= prm.tmaEnd || start >= totalBytes) return;
    const unsigned avail = totalBytes - start < blockBytes ? static_cast<unsigned>(totalBytes - start) : blockBytes;
    const unsigned from = prm.tmaEnd > start ? static_cast<unsigned>(prm.tmaEnd - start) : 0u;
    unsigned char* dst = slots + (warp * S + slot) * slotBytes;
    for (unsigned b = from + lane; b < avail; b += 32u) dst[b ^ ((b >> 3) & prm.swzMask)] = gin[start + b];
    __syncwarp();
  };
  // (lane 0) pull the block of tile t into L2: the ring holds S blocks per warp, L2 extends the pipeline past that at no
  // cost in shared memory
  auto prefetchBlock = [&](unsigned t) {
    if (lane == 0 && t < nTiles) {
      const unsigned long long start = blockStart(t, warp);
      const int c1 = static_cast<int>(start >> prm.wShift), c2 = static_cast<int>((start & ((1u << prm.wShift) - 1u)) >> 4);
      asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(&tmap), "r"(0), "r"(c1), "r"(c2) : "memory");
    }
  };
  if (warp < NW) {
    for (unsigned t = 0; t < S && t < nTiles; t++) issueBlock(t, t);
    for (unsigned t = S; t < S + prm.prefetch; t++) prefetchBlock(t);
  }

  // carry / done bookkeeping is a pure function of the tile index: every warp tracks it on its own.  `carry` = demod samples
  // before the start of the tile that belong to windows not yet finished; `done` = audio outputs produced by earlier tiles.
  unsigned carry = 0;
  unsigned long long done = 0;
  const unsigned otDiv = OT / D2, otRem = OT % D2;
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

  const bool pairs = (D2 & 1u) == 0;
  mbarWait(constBar, 0);
  if (warp < NW) {
  // =========================== compute warps ===========================
  const unsigned nFull = prm.KS >> 1, odd = prm.KS & 1u, Q = prm.Q;
  const uint4* bLane = bFrag + lane * 3u;
  unsigned slot = 0, slotPhase = 0;
  for (unsigned t = 0; t < nTiles; t++) {
    mbarWait(&full[warp * S + slot], slotPhase);
    patchBlock(t, slot);
    // ldmatrix row offset of this lane in the (unswizzled) block: matrix mi = lane / 8 -> rows +8 for odd mi, k +16 for mi >= 2
    const uint32_t slotBase = smemAddr(slots + (warp * S + slot) * slotBytes);
    const unsigned laneOff = ((lane & 7u) + 8u * ((lane >> 3) & 1u)) * AS + 16u * (lane >> 4);

    int acc[G][3][4];
#pragma unroll
    for (int j = 0; j < G; j++)
#pragma unroll
      for (int d = 0; d < 3; d++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[j][d][e] = MAGIC ? static_cast<int>(kToepMagicBits) : 0;

    // k-step pairs are loaded in order; pair q covers bytes [64q, 64q + 64) of the A-row.  m-tiles are 16*AS = 128*D bytes
    // apart, a multiple of 1024, so they share the swizzle term.
    unsigned ldQ = 0;
    auto loadNext = [&](ToepFrag<G>& f) {
      const unsigned o0 = laneOff + ldQ * 64u, o1 = o0 + 32u;
      const uint32_t p0 = slotBase + (o0 ^ ((o0 >> 3) & prm.swzMask)), p1 = slotBase + (o1 ^ ((o1 >> 3) & prm.swzMask));
#pragma unroll
      for (int j = 0; j < G; j++) {
        ldsm4(f.a[j][0], p0 + j * 16u * AS);
        ldsm4(f.a[j][1], p1 + j * 16u * AS);
      }
#pragma unroll
      for (int i = 0; i < 3; i++) f.b[i] = bLane[ldQ * 96u + i];
      ldQ++;
    };
    ToepFrag<G> f0, f1;
    loadNext(f0);
    unsigned q = 0;
#pragma unroll 1
    while (q + 2 <= nFull) {
      loadNext(f1);
      toepMma<G>(acc, f0, true);
      if (ldQ < Q) loadNext(f0);
      toepMma<G>(acc, f1, true);
      q += 2;
    }
    if (q < nFull) {
      if (odd) loadNext(f1);
      toepMma<G>(acc, f0, true);
      if (odd) toepMma<G>(acc, f1, false);
    } else if (odd) {
      toepMma<G>(acc, f0, false);
    }

    // ---- the slot has been consumed (every load fed an IMMA that has issued): refill it -----------------------
    if (t + S < nTiles) issueBlock(t + S, slot);
    if (prm.prefetch) prefetchBlock(t + S + prm.prefetch);

    // ---- digits -> float: outputs lane and 32 + lane of each m-tile --------------------------------------------
    float2 y[G][2];
#pragma unroll
    for (int j = 0; j < G; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        float2 v0, v1, v2;
        if constexpr (MAGIC) {
          const float2 mC = make_float2(-kToepMagic, -kToepMagic);
          v0 = __fadd2_rn(make_float2(__int_as_float(acc[j][0][2 * h]), __int_as_float(acc[j][0][2 * h + 1])), mC);
          v1 = __fadd2_rn(make_float2(__int_as_float(acc[j][1][2 * h]), __int_as_float(acc[j][1][2 * h + 1])), mC);
          v2 = __fadd2_rn(make_float2(__int_as_float(acc[j][2][2 * h]), __int_as_float(acc[j][2][2 * h + 1])), mC);
        } else {
          v0 = make_float2(static_cast<float>(acc[j][0][2 * h]), static_cast<float>(acc[j][0][2 * h + 1]));
          v1 = make_float2(static_cast<float>(acc[j][1][2 * h]), static_cast<float>(acc[j][1][2 * h + 1]));
          v2 = make_float2(static_cast<float>(acc[j][2][2 * h]), static_cast<float>(acc[j][2][2 * h + 1]));
        }
        y[j][h] = axpy2(prm.s2, v2, axpy2(prm.s1, v1, scale2(prm.s0, v0)));
      }

    // ---- demodulate ------------------------------------------------------------------------------------------
    float dmv[G][2];
    if (fm) {
#pragma unroll
      for (int j = 0; j < G; j++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          // successor of output (j, h, lane): lane + 1 of the same half, or lane 0 of the next half
          const float2 c = y[j][h];
          float2 n;
          n.x = __shfl_down_sync(0xffffffffu, c.x, 1);
          n.y = __shfl_down_sync(0xffffffffu, c.y, 1);
          float2 first = make_float2(0.0f, 0.0f);
          if (h == 0) {
            first.x = __shfl_sync(0xffffffffu, y[j][1].x, 0);
            first.y = __shfl_sync(0xffffffffu, y[j][1].y, 0);
          } else if (j + 1 < G) {
            first.x = __shfl_sync(0xffffffffu, y[j + 1 < G ? j + 1 : j][0].x, 0);
            first.y = __shfl_sync(0xffffffffu, y[j + 1 < G ? j + 1 : j][0].y, 0);
          }
          if (lane == 31u) n = first;
          const float2 d = make_float2(fmaf(n.y, c.y, n.x * c.x), fmaf(n.y, c.x, -n.x * c.y));
          const float2 r = cmulf(d, prm.rot1);
          dmv[j][h] = prm.gain * atan2f(r.y, r.x);
        }
    } else {
#pragma unroll
      for (int j = 0; j < G; j++)
#pragma unroll
        for (int h = 0; h < 2; h++) dmv[j][h] = sqrtf(fmaf(y[j][h].x, y[j][h].x, y[j][h].y * y[j][h].y));
    }

    // ---- into the demod ring, once no unfinished audio window reaches into the tile that lived there ------------
    const unsigned line = t & (kToepLines - 1u), use = t / kToepLines;
    if (use > 0) mbarWait(&tileDone[line], (use - 1u) & 1u);
    float* dst = ring + line * OT + warp * OTW;
#pragma unroll
    for (int j = 0; j < G; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const unsigned o = j * 64u + h * 32u + lane;
        if (o < OTW) {
          dst[o] = dmv[j][h];
          const unsigned idx = line * OT + warp * OTW + o;      // position in the ring
          if (idx < mirror) ring[R + idx] = dmv[j][h];          // windows that cross the end of the ring read on here
        }
      }
    __syncwarp();
    if (lane == 0) mbarArrive(&tileFull[line]);
    if (++slot == S) {
      slot = 0;
      slotPhase ^= 1u;
    }
  }
  } else {
  // =========================== audio warps: tile t belongs to audio warp t mod NA ===========================
  const unsigned me = warp - NW;
  unsigned mine = 0;  // t mod NA
  for (unsigned t = 0; t < nTiles; t++) {
    const unsigned line = t & (kToepLines - 1u), use = t / kToepLines;
    const unsigned nA = outputsReady(carry);
    if (mine == me) {
      mbarWait(&tileFull[line], use & 1u);
      unsigned base = line * OT + R - carry;  // ring index of the first unfinished window
      if (base >= R) base -= R;
      // each lane works on outputs o and o + 32 at once: two independent dot products
      for (unsigned o = lane; o < nA; o += 64u) {
        const bool two = o + 32u < nA;
        unsigned sa = base + o * D2, sb = two ? sa + 32u * D2 : sa;
        if (sa >= R) sa -= R;
        if (sb >= R) sb -= R;
        if (sb >= R) sb -= R;
        const float* xa = ring + sa;
        const float* xb = ring + sb;
        float a0s = 0.0f, a1s = 0.0f, a2s = 0.0f, a3s = 0.0f, b0s = 0.0f, b1s = 0.0f, b2s = 0.0f, b3s = 0.0f;
        unsigned j = 0;
        if (pairs) {  // window starts are even: 64-bit loads
          // 16 taps per iteration, every load issued before the first multiply-add (shared-memory latency under load is
          // ~100 cycles)
#pragma unroll 1
          for (; j + 16 <= T2; j += 16) {
            float4 hv[4];
            float2 p[4][2], qq[4][2];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              hv[u] = *reinterpret_cast<const float4*>(h2 + j + 4 * u);
              p[u][0] = *reinterpret_cast<const float2*>(xa + j + 4 * u);
              p[u][1] = *reinterpret_cast<const float2*>(xa + j + 4 * u + 2);
              qq[u][0] = *reinterpret_cast<const float2*>(xb + j + 4 * u);
              qq[u][1] = *reinterpret_cast<const float2*>(xb + j + 4 * u + 2);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              a0s = fmaf(hv[u].x, p[u][0].x, a0s);
              a1s = fmaf(hv[u].y, p[u][0].y, a1s);
              a2s = fmaf(hv[u].z, p[u][1].x, a2s);
              a3s = fmaf(hv[u].w, p[u][1].y, a3s);
              b0s = fmaf(hv[u].x, qq[u][0].x, b0s);
              b1s = fmaf(hv[u].y, qq[u][0].y, b1s);
              b2s = fmaf(hv[u].z, qq[u][1].x, b2s);
              b3s = fmaf(hv[u].w, qq[u][1].y, b3s);
            }
          }
        } else {
#pragma unroll 1
          for (; j + 8 <= T2; j += 8) {
            float4 hv[2];
            float xs[8], zs[8];
#pragma unroll
            for (int u = 0; u < 2; u++) hv[u] = *reinterpret_cast<const float4*>(h2 + j + 4 * u);
#pragma unroll
            for (int u = 0; u < 8; u++) {
              xs[u] = xa[j + u];
              zs[u] = xb[j + u];
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
              a0s = fmaf(hv[u].x, xs[4 * u], a0s);
              a1s = fmaf(hv[u].y, xs[4 * u + 1], a1s);
              a2s = fmaf(hv[u].z, xs[4 * u + 2], a2s);
              a3s = fmaf(hv[u].w, xs[4 * u + 3], a3s);
              b0s = fmaf(hv[u].x, zs[4 * u], b0s);
              b1s = fmaf(hv[u].y, zs[4 * u + 1], b1s);
              b2s = fmaf(hv[u].z, zs[4 * u + 2], b2s);
              b3s = fmaf(hv[u].w, zs[4 * u + 3], b3s);
            }
          }
        }
        for (; j < T2; j++) {
          a0s = fmaf(h2[j], xa[j], a0s);
          b0s = fmaf(h2[j], xb[j], b0s);
        }
        prm.out[a0 + done + o] = (a0s + a1s) + (a2s + a3s);
        if (two) prm.out[a0 + done + o + 32u] = (b0s + b1s) + (b2s + b3s);
      }
      __syncwarp();
      // tiles t - span .. t are not needed by this batch any more (each waits for span + 1 such releases)
      if (lane == 0) {
        for (unsigned i = 0; i <= prm.span && i <= t; i++) mbarArrive(&tileDone[(t - i) & (kToepLines - 1u)]);
      }
    }
    done += nA;
    carry = carry + OT - nA * D2;
    if (++mine == NA) mine = 0;
  }
  }
}

#endif  // __CUDACC__

}  // namespace b200sdr
