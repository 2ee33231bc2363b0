// windowKernel -- register-tiled decimating FIR for the FFMA-bound cells of the roofline sweep (many taps per kept
// output: ceil(T / D) > 8), real taps x complex float data (gsdrFirFC) and real x real (gsdrFirFF).
//
// Polyphase view: y[k] = sum_p sum_m h[m*D + p] * x[(k + m)*D + p].  For a fixed phase p this is an ordinary (stride-1)
// FIR of length M = ceil(T/D) over the decimated sequence x_p[q] = x[q*D + p], so one thread that owns R CONSECUTIVE
// outputs can slide a window of R samples through its registers: per tap one new sample is read from shared memory and
// R FMAs (packed FFMA2 for complex data) are issued -- an FMA : load ratio of R : 1 instead of the 1 : 1 of a
// thread-per-output kernel.  The input is staged in shared memory one (phase chunk, tap chunk) at a time, transposed to
// [phase][q] with one element of padding per R so that the R-strided reads of a warp are conflict-free; every input
// sample of the CTA's window is read from HBM once.
#include <gsdr/gsdr.h>

#include <cstdlib>

#include "fir_dispatch.h"
#include "fir_kernels.cuh"

namespace b200sdr {

namespace {

constexpr int kWinThreads = 128;
constexpr int kWinR = 8;          // consecutive outputs per thread (16 measured slower for real data: 0.30 vs 0.24 ms on the C5 audio stage -- fewer resident CTAs)
constexpr int kWinMaxTapChunk = 64;  // taps (per phase) staged at a time: ceil(T/D) rounded up to 8, at most this many.  The whole
                                     // window of a chunk (outputs + taps) is staged per chunk, so a chunk as long as the filter
                                     // (M <= 64: the audio FIRs, the T/D ~ 16 cells of the sweep) stages every sample ONCE and
                                     // multiplies no padding zeros; long filters stage half as often as with 32-tap chunks
// phases staged together: the largest of 5, 4, 2, 1 that divides D (5: the audio decimations 5 and 10 -- one coalesced pass
// instead of five strided ones; 8 measured slower: the 76 KB tile halves the resident CTAs)

template <typename T>
struct WinTraits;
template <>
struct WinTraits<float2> {
  using Vec = float4;  // two elements
  __device__ static void unpack(float4 v, float2 (&x)[2]) {
    x[0] = make_float2(v.x, v.y);
    x[1] = make_float2(v.z, v.w);
  }
  __device__ static float2 zero() { return make_float2(0.0f, 0.0f); }
  __device__ static float2 fma(float h, float2 x, float2 acc) { return __ffma2_rn(make_float2(h, h), x, acc); }
};
template <>
struct WinTraits<float> {
  using Vec = float4;  // four elements
  __device__ static void unpack(float4 v, float (&x)[4]) {
    x[0] = v.x;
    x[1] = v.y;
    x[2] = v.z;
    x[3] = v.w;
  }
  __device__ static float zero() { return 0.0f; }
  __device__ static float fma(float h, float x, float acc) { return fmaf(h, x, acc); }
};

// Where a CTA's samples come from.  Plain: one stream of Elem.  Pair: TWO real streams (two channels of the channelizer's audio
// stage, `stride` floats apart) travel through the kernel as the two halves of a float2, so the real-data FIR runs on packed FFMA2
// like the complex one: with scalar FFMA every load and address instruction takes an issue slot from the FMA pipe, with FFMA2 the
// pipe is busy two cycles per issue and the rest fits in between.
template <typename Elem, bool PAIR>
struct WinSource {
  static constexpr unsigned VEC = 16 / sizeof(Elem);
  const Elem* p;
  __device__ WinSource(const FirParams& prm, unsigned y) : p(static_cast<const Elem*>(prm.in) + y * prm.inBatchStride) {}
  __device__ Elem load(unsigned long long i) const { return p[i]; }
  __device__ void loadVec(unsigned long long i, Elem (&x)[VEC]) const { WinTraits<Elem>::unpack(*reinterpret_cast<const float4*>(p + i), x); }
  __device__ unsigned misalign(unsigned long long i) const { return static_cast<unsigned>(reinterpret_cast<uintptr_t>(p + i) / sizeof(Elem)) % VEC; }
  __device__ bool vecOk() const { return true; }
};
template <>
struct WinSource<float2, true> {
  static constexpr unsigned VEC = 4;
  const float *a, *b;  // b == nullptr: an odd number of streams, the last one travels alone
  __device__ WinSource(const FirParams& prm, unsigned y)
      : a(static_cast<const float*>(prm.in) + 2ull * y * prm.inBatchStride), b(2 * y + 1 < prm.winStreams ? a + prm.inBatchStride : nullptr) {}
  __device__ float2 load(unsigned long long i) const { return make_float2(a[i], b ? b[i] : 0.0f); }
  __device__ void loadVec(unsigned long long i, float2 (&x)[4]) const {
    const float4 u = *reinterpret_cast<const float4*>(a + i);
    const float4 v = b ? *reinterpret_cast<const float4*>(b + i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    x[0] = make_float2(u.x, v.x);
    x[1] = make_float2(u.y, v.y);
    x[2] = make_float2(u.z, v.z);
    x[3] = make_float2(u.w, v.w);
  }
  __device__ unsigned misalign(unsigned long long i) const { return static_cast<unsigned>(reinterpret_cast<uintptr_t>(a + i) / sizeof(float)) % 4u; }
  __device__ bool vecOk() const { return b == nullptr || (reinterpret_cast<uintptr_t>(a) - reinterpret_cast<uintptr_t>(b)) % 16 == 0; }
};

template <int R>
__host__ __device__ constexpr unsigned winPadded(unsigned q) { return q + q / R; }  // one pad element per R

// Staged samples keep the order they have in memory: sample (q, pc) -- decimated index q, phase pc of the PC staged -- sits at
// element q * PC + pc + q / R.  The pad element per R decimated samples makes the stride between the windows of neighbouring
// threads odd (R * PC + 1), so the R-strided reads of a warp are conflict-free, and for D == PC the whole tile is ONE contiguous
// run of the input: it is copied with 16-byte loads.
template <int PC, int R>
__host__ __device__ constexpr unsigned winSlot(unsigned e) { return e + e / (PC * R); }  // e = q * PC + pc

// PC = phases staged together (divides D); TC = taps per phase staged together (a multiple of R, prm.winTapChunk).
// Shared memory: taps [PC][kWinMaxTapChunk] floats, then the staged samples.
template <typename Elem, int PC, int kWinR, bool PAIR>
__global__ void __launch_bounds__(kWinThreads) windowKernel(const FirParams prm) {
  extern __shared__ __align__(16) unsigned char wsmem[];
  using Source = WinSource<Elem, PAIR>;
  constexpr unsigned BO = kWinThreads * kWinR;      // outputs per CTA
  constexpr unsigned VEC = Source::VEC;             // elements per vector load (16 bytes of every stream)
  const unsigned TC = prm.winTapChunk;
  const unsigned ROWS = BO + TC;                    // decimated samples staged per phase (window of the tap chunk)
  const unsigned total = ROWS * PC;                 // elements staged per pass
  float* sTaps = reinterpret_cast<float*>(wsmem);
  Elem* sX = reinterpret_cast<Elem*>(wsmem + PC * kWinMaxTapChunk * sizeof(float));

  const unsigned tid = threadIdx.x;
  const unsigned D = prm.D, T = prm.T, M = prm.M;
  const unsigned long long k0 = static_cast<unsigned long long>(blockIdx.x) * BO;  // first output of the CTA
  const Source src(prm, blockIdx.y);  // blockIdx.y: independent streams (batched), or pairs of them
  const bool contiguous = D == PC;
  // strided tiles: every decimated sample contributes a run of PC elements; whole 16-byte pieces when everything is aligned
  const bool stridedVec = !contiguous && PC % VEC == 0 && D % VEC == 0 && src.vecOk() && src.misalign(0) == 0;

  Elem acc[kWinR];
#pragma unroll
  for (int r = 0; r < kWinR; r++) acc[r] = WinTraits<Elem>::zero();

  for (unsigned p0 = 0; p0 < D; p0 += PC) {
    for (unsigned m0 = 0; m0 < M; m0 += TC) {
      __syncthreads();  // the previous chunk has been consumed
      // taps of this chunk: sTaps[pc][mm] = h[(m0+mm)*D + p0+pc]
      for (unsigned i = tid; i < PC * TC; i += kWinThreads) {
        const unsigned pc = i / TC, mm = i % TC;
        const unsigned long long j = static_cast<unsigned long long>(m0 + mm) * D + p0 + pc;
        sTaps[i] = (m0 + mm < M && j < T) ? prm.taps[j] : 0.0f;
      }
      // samples: element e = q * PC + pc of the tile is x[(k0 + m0 + q) * D + p0 + pc]
      const unsigned long long base = (k0 + m0) * D + p0;  // input index of element 0
      if (contiguous) {
        const unsigned long long left = base < prm.nIn ? prm.nIn - base : 0;
        const unsigned valid = left < total ? static_cast<unsigned>(left) : total;  // elements past it are zero
        unsigned head = total, nvec = 0;  // [0, head) one by one, nvec vectors, the rest one by one
        if (src.vecOk()) {
          head = (VEC - src.misalign(base)) % VEC;
          if (head > total) head = total;
          nvec = (total - head) / VEC;
        }
        for (unsigned e = tid; e < head; e += kWinThreads) sX[winSlot<PC, kWinR>(e)] = e < valid ? src.load(base + e) : WinTraits<Elem>::zero();
#pragma unroll 4
        for (unsigned v = tid; v < nvec; v += kWinThreads) {
          const unsigned e = head + v * VEC;
          Elem x[VEC];
          if (e + VEC <= valid) {
            src.loadVec(base + e, x);
          } else {
#pragma unroll
            for (unsigned i = 0; i < VEC; i++) x[i] = e + i < valid ? src.load(base + e + i) : WinTraits<Elem>::zero();
          }
#pragma unroll
          for (unsigned i = 0; i < VEC; i++) sX[winSlot<PC, kWinR>(e + i)] = x[i];
        }
        for (unsigned e = head + nvec * VEC + tid; e < total; e += kWinThreads)
          sX[winSlot<PC, kWinR>(e)] = e < valid ? src.load(base + e) : WinTraits<Elem>::zero();
      } else if (stridedVec) {
#pragma unroll 4
        for (unsigned v = tid; v < total / VEC; v += kWinThreads) {
          const unsigned e = v * VEC, q = e / PC, pc = e % PC;
          const unsigned long long idx = base + static_cast<unsigned long long>(q) * D + pc;
          Elem x[VEC];
          if (idx + VEC <= prm.nIn) {
            src.loadVec(idx, x);
          } else {
#pragma unroll
            for (unsigned i = 0; i < VEC; i++) x[i] = idx + i < prm.nIn ? src.load(idx + i) : WinTraits<Elem>::zero();
          }
#pragma unroll
          for (unsigned i = 0; i < VEC; i++) sX[winSlot<PC, kWinR>(e + i)] = x[i];
        }
      } else {
        for (unsigned e = tid; e < total; e += kWinThreads) {
          const unsigned q = e / PC, pc = e % PC;
          const unsigned long long idx = base + static_cast<unsigned long long>(q) * D + pc;
          sX[winSlot<PC, kWinR>(e)] = idx < prm.nIn ? src.load(idx) : WinTraits<Elem>::zero();
        }
      }
      __syncthreads();

#pragma unroll
      for (int pc = 0; pc < PC; pc++) {
        const float* hs = sTaps + pc * TC;
        // this thread's window starts at decimated sample R * tid of the tile: element tid * (R * PC + 1) + pc; sample j of the
        // window (j = R a + b) sits (R a + b) * PC + a elements further -- compile-time constants off `xw` inside a block of R taps
        const Elem* xw = sX + tid * (kWinR * PC + 1) + pc;
        Elem w[kWinR];
#pragma unroll
        for (int r = 0; r < kWinR - 1; r++) w[r] = xw[r * PC];
#pragma unroll 2
        for (unsigned mm = 0; mm < TC; mm += kWinR, xw += kWinR * PC + 1) {
          // eight taps at a time (TC is a multiple of 8); before tap tau the window holds samples base + tau .. + R - 1, sample
          // base + tau + i in slot (tau + i) % R -- all slot numbers below are compile-time constants
#pragma unroll
          for (int blk = 0; blk < kWinR; blk += 8) {
            if (mm + blk < TC) {
              const float4 h0 = *reinterpret_cast<const float4*>(hs + mm + blk), h1 = *reinterpret_cast<const float4*>(hs + mm + blk + 4);
              const float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
              for (int u = 0; u < 8; u++) {
                constexpr int kLast = kWinR - 1;
                const int c = blk + u + kLast;  // newest sample of the window, relative to the R-aligned block start
                w[c % kWinR] = xw[c * PC + c / kWinR];
#pragma unroll
                for (int r = 0; r < kWinR; r++) acc[r] = WinTraits<Elem>::fma(h[u], w[(blk + u + r) % kWinR], acc[r]);
              }
            }
          }
        }
      }
    }
  }

  // Outputs leave through shared memory so that every store instruction of a warp covers one contiguous run (thread-major
  // stores would touch 4 bytes of 32 different sectors per instruction: harmless in the local L2, which merges them, but each of
  // them is a separate small write when `out` is a peer GPU's memory over NVLink -- the multi-GPU channelizer).
  __syncthreads();
  Elem* sOut = sX;  // BO + BO / R elements <= one staged row
#pragma unroll
  for (int r = 0; r < kWinR; r++) sOut[tid * (kWinR + 1) + r] = acc[r];
  __syncthreads();
  if constexpr (PAIR) {
    float* outA = static_cast<float*>(prm.out) + 2ull * blockIdx.y * prm.outBatchStride;
    float* outB = 2 * blockIdx.y + 1 < prm.winStreams ? outA + prm.outBatchStride : nullptr;
#pragma unroll
    for (int j = 0; j < kWinR; j++) {
      const unsigned o = j * kWinThreads + tid;
      const unsigned long long k = k0 + o;
      if (k < prm.nOut) {
        const float2 y = sOut[winPadded<kWinR>(o)];
        outA[k] = y.x;
        if (outB) outB[k] = y.y;
      }
    }
  } else {
    Elem* out = static_cast<Elem*>(prm.out) + blockIdx.y * prm.outBatchStride;
#pragma unroll
    for (int j = 0; j < kWinR; j++) {
      const unsigned o = j * kWinThreads + tid;
      const unsigned long long k = k0 + o;
      if (k < prm.nOut) out[k] = sOut[winPadded<kWinR>(o)];
    }
  }
}

template <typename Elem, int kWinR, bool PAIR>
cudaError_t launchWindowT(FirParams prm, unsigned batch, cudaStream_t stream) {
  const unsigned pc = prm.D % 5 == 0 ? 5u : prm.D % 4 == 0 ? 4u : prm.D % 2 == 0 ? 2u : 1u;
  constexpr unsigned BO = kWinThreads * kWinR;
  unsigned tc = (prm.M + 7u) / 8u * 8u;
  if (tc > static_cast<unsigned>(kWinMaxTapChunk)) tc = kWinMaxTapChunk;
  prm.winTapChunk = tc;
  const unsigned ROWS = BO + tc;
  const size_t smem = pc * kWinMaxTapChunk * sizeof(float) + static_cast<size_t>(ROWS * pc + ROWS / kWinR + 1) * sizeof(Elem);
  const unsigned long long blocks = (prm.nOut + BO - 1) / BO;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  void (*k)(const FirParams) = pc == 5   ? windowKernel<Elem, 5, kWinR, PAIR>
                               : pc == 4 ? windowKernel<Elem, 4, kWinR, PAIR>
                               : pc == 2 ? windowKernel<Elem, 2, kWinR, PAIR>
                                         : windowKernel<Elem, 1, kWinR, PAIR>;
  if (smem > 48 * 1024) {
    const cudaError_t e = ensureDynamicSmem(reinterpret_cast<const void*>(k), 100 * 1024);
    if (e != cudaSuccess) return e;
  }
  if (batch == 0 || batch > 65535u) return cudaErrorInvalidConfiguration;
  prm.winStreams = batch;
  k<<<dim3(static_cast<unsigned>(blocks), PAIR ? (batch + 1) / 2 : batch), kWinThreads, smem, stream>>>(prm);
  return launchStatus();
}

}  // namespace

bool windowEligible(int elem, bool tapsComplex, bool mix, const FirParams& prm) {
  if (tapsComplex || mix || prm.mod != kModNone) return false;
  if (elem != kElemComplex && elem != kElemReal) return false;
  return prm.D > 0 && (prm.T + prm.D - 1) / prm.D > 8;  // otherwise the rows / staged direct kernels are the better shape
}

cudaError_t launchWindow(int elem, FirParams prm, cudaStream_t stream) { return launchWindowBatched(elem, prm, 1, stream); }

// `batch` independent streams, prm.inBatchStride / prm.outBatchStride elements apart (the channelizer's audio stage)
cudaError_t launchWindowBatched(int elem, FirParams prm, unsigned batch, cudaStream_t stream) {
  prm.M = (prm.T + prm.D - 1) / prm.D;
  if (batch == 1) prm.inBatchStride = prm.outBatchStride = 0;
  if (elem == kElemComplex) return launchWindowT<float2, kWinR, false>(prm, batch, stream);
  // real data: two streams at a time as the halves of a float2 (see WinSource); a single stream keeps the scalar kernel
  static const bool noPair = std::getenv("B200SDR_WINDOW_NO_PAIR") != nullptr;
  if (batch >= 2 && !noPair) return launchWindowT<float2, kWinR, true>(prm, batch, stream);
  return launchWindowT<float, kWinR, false>(prm, batch, stream);
}

}  // namespace b200sdr
