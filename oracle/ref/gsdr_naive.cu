// gsdr_naive.cu -- a deliberately straightforward restatement of the `gsdr` kernel library the reference
// links against (github.com/kernrj/gsdr @ main, un-vendored and un-pinned: /root/reference/src/CMakeLists.txt:13-19;
// its source is absent from /root/reference and from this image).
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It exists so that the reference's own host framework, compiled in
// place from /root/reference by build_ref.sh, can run on a GPU as the "reference CUDA pipeline" baseline (B1 in
// SURVEY.md section 8(d)) and as an independent second implementation for parity tests.  One thread per output
// element, plain global loads, no shared memory, no vectorisation, fp32 arithmetic: what a first CUDA port of
// each op looks like.  Signatures are reconstructed from the call sites cited at each function.
#include <cuComplex.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#define NAIVE_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

constexpr unsigned kThreads = 256;

inline unsigned blocksFor(size_t n) { return static_cast<unsigned>((n + kThreads - 1) / kThreads); }

struct DevicePush {
  int prev = -1;
  explicit DevicePush(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DevicePush() {
    int now = -1;
    if (cudaGetDevice(&now) == cudaSuccess && now != prev && prev >= 0) cudaSetDevice(prev);
  }
};

__global__ void int8ToNormFloat(const int8_t* in, float* out, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(in[i]) * (1.0f / 128.0f);
}

// phi_i = phiStart + i*(phiEnd - phiStart)/n as an exact function of the two float32 arguments (fp64 arithmetic), so
// that the restated kernels add no phase error of their own to the float32 phase the reference's host code tracks
__device__ double phaseOf(float phiStart, float phiEnd, size_t i, size_t n) {
  return static_cast<double>(phiStart) +
         static_cast<double>(i) * ((static_cast<double>(phiEnd) - static_cast<double>(phiStart)) / static_cast<double>(n));
}

__global__ void cosineF(float phiStart, float phiEnd, float* out, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(cos(phaseOf(phiStart, phiEnd, i, n)));
}

__global__ void cosineC(float phiStart, float phiEnd, cuComplex* out, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    const double phi = phaseOf(phiStart, phiEnd, i, n);
    out[i] = make_cuComplex(static_cast<float>(cos(phi)), static_cast<float>(sin(phi)));
  }
}

__global__ void multiplyCC(const cuComplex* a, const cuComplex* b, cuComplex* out, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = cuCmulf(a[i], b[i]);
}

template <typename Tap, typename In, typename Out>
__device__ Out macc(Out acc, Tap h, In x);
template <>
__device__ float macc<float, float, float>(float acc, float h, float x) { return fmaf(h, x, acc); }
template <>
__device__ cuComplex macc<float, cuComplex, cuComplex>(cuComplex acc, float h, cuComplex x) {
  return make_cuComplex(fmaf(h, x.x, acc.x), fmaf(h, x.y, acc.y));
}
template <>
__device__ cuComplex macc<cuComplex, cuComplex, cuComplex>(cuComplex acc, cuComplex h, cuComplex x) {
  return cuCaddf(acc, cuCmulf(h, x));
}
template <>
__device__ cuComplex macc<cuComplex, float, cuComplex>(cuComplex acc, cuComplex h, float x) {
  return make_cuComplex(fmaf(h.x, x, acc.x), fmaf(h.y, x, acc.y));
}

template <typename T>
__device__ T zero();
template <>
__device__ float zero<float>() { return 0.0f; }
template <>
__device__ cuComplex zero<cuComplex>() { return make_cuComplex(0.0f, 0.0f); }

// out[k] = sum_j taps[j] * in[k*D + j]   (correlation order; /root/reference/src/filters/Fir.cpp:124 passes
// the taps as given and names them "tapsReversed"; pinned by /root/reference/tests/FirTests.cpp)
template <typename Tap, typename In, typename Out>
__global__ void fir(size_t D, const Tap* taps, size_t T, const In* in, Out* out, size_t nOut) {
  const size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (k >= nOut) return;
  const In* x = in + k * D;
  Out acc = zero<Out>();
  for (size_t j = 0; j < T; j++) acc = macc<Tap, In, Out>(acc, taps[j], x[j]);
  out[k] = acc;
}

__global__ void magnitude(const cuComplex* in, float* out, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = sqrtf(in[i].x * in[i].x + in[i].y * in[i].y);
}

__global__ void quadFmDemod(const cuComplex* in, float* out, float gain, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    const cuComplex d = cuCmulf(in[i + 1], cuConjf(in[i]));
    out[i] = gain * atan2f(d.y, d.x);
  }
}

__global__ void addConst(const float* in, float c, float* out, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = in[i] + c;
}

__global__ void addToMagnitude(const cuComplex* in, float c, cuComplex* out, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    const float m = sqrtf(in[i].x * in[i].x + in[i].y * in[i].y);
    const float s = (m + c) / m;
    out[i] = make_cuComplex(in[i].x * s, in[i].y * s);
  }
}

// The upstream author's fused op (/root/reference/src/applications/fm_simpletest.cpp:400-413), restated as
// its unfused definition: mix, low-pass + decimate, discriminate.  One thread per output, everything recomputed.
__global__ void fmDemod(float phaseStepRad, size_t D, size_t firstSampleOffset, const float* taps, size_t T,
                        const cuComplex* in, float* out, float gain, size_t nOut) {
  const size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (k >= nOut) return;
  cuComplex y[2];
  for (int s = 0; s < 2; s++) {
    const size_t base = (k + s) * D;
    cuComplex acc = make_cuComplex(0.0f, 0.0f);
    for (size_t j = 0; j < T; j++) {
      const double phi = static_cast<double>(phaseStepRad) * static_cast<double>(firstSampleOffset + base + j);
      const cuComplex w = make_cuComplex(static_cast<float>(cos(phi)), static_cast<float>(sin(phi)));
      const cuComplex z = cuCmulf(in[base + j], w);
      acc.x = fmaf(taps[j], z.x, acc.x);
      acc.y = fmaf(taps[j], z.y, acc.y);
    }
    y[s] = acc;
  }
  const cuComplex d = cuCmulf(y[1], cuConjf(y[0]));
  out[k] = gain * atan2f(d.y, d.x);
}

}  // namespace

#define LAUNCH(kernel, n, stream, ...)                                   \
  do {                                                                   \
    DevicePush push__(cudaDevice);                                       \
    if ((n) > 0) kernel<<<blocksFor(n), kThreads, 0, stream>>>(__VA_ARGS__); \
    return cudaGetLastError();                                           \
  } while (false)

// /root/reference/src/filters/Int8ToFloat.cpp:89-94
NAIVE_EXPORT cudaError_t gsdrInt8ToNormFloat(const int8_t* in, float* out, size_t n, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH(int8ToNormFloat, n, s, in, out, n);
}
// /root/reference/src/filters/CosineSource.cpp:74-80
NAIVE_EXPORT cudaError_t gsdrCosineF(float phiStart, float phiEnd, float* out, size_t n, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH(cosineF, n, s, phiStart, phiEnd, out, n);
}
// /root/reference/src/filters/ComplexCosineSource.cpp:74-80
NAIVE_EXPORT cudaError_t gsdrCosineC(float phiStart, float phiEnd, cuComplex* out, size_t n, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH(cosineC, n, s, phiStart, phiEnd, out, n);
}
// /root/reference/src/filters/Multiply.cpp:145-151
NAIVE_EXPORT cudaError_t gsdrMultiplyCC(const cuComplex* a, const cuComplex* b, cuComplex* out, size_t n, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH(multiplyCC, n, s, a, b, out, n);
}
// /root/reference/src/filters/Fir.cpp:230-268
NAIVE_EXPORT cudaError_t gsdrFirFF(size_t D, const float* taps, size_t T, const float* in, float* out, size_t nOut, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH((fir<float, float, float>), nOut, s, D, taps, T, in, out, nOut);
}
NAIVE_EXPORT cudaError_t gsdrFirFC(size_t D, const float* taps, size_t T, const cuComplex* in, cuComplex* out, size_t nOut, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH((fir<float, cuComplex, cuComplex>), nOut, s, D, taps, T, in, out, nOut);
}
NAIVE_EXPORT cudaError_t gsdrFirCC(size_t D, const cuComplex* taps, size_t T, const cuComplex* in, cuComplex* out, size_t nOut, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH((fir<cuComplex, cuComplex, cuComplex>), nOut, s, D, taps, T, in, out, nOut);
}
NAIVE_EXPORT cudaError_t gsdrFirCF(size_t D, const cuComplex* taps, size_t T, const float* in, cuComplex* out, size_t nOut, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH((fir<cuComplex, float, cuComplex>), nOut, s, D, taps, T, in, out, nOut);
}
// /root/reference/src/filters/QuadAmDemod.cpp:93-98, Magnitude.cpp:91-96
NAIVE_EXPORT cudaError_t gsdrQuadAmDemod(const cuComplex* in, float* out, size_t n, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH(magnitude, n, s, in, out, n);
}
NAIVE_EXPORT cudaError_t gsdrMagnitude(const cuComplex* in, float* out, size_t n, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH(magnitude, n, s, in, out, n);
}
// /root/reference/src/filters/QuadFmDemod.cpp:98-104
NAIVE_EXPORT cudaError_t gsdrQuadFmDemod(const cuComplex* in, float* out, float gain, size_t n, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH(quadFmDemod, n, s, in, out, gain, n);
}
// /root/reference/src/filters/AddConst.cpp:99-105
NAIVE_EXPORT cudaError_t gsdrAddConstFF(const float* in, float c, float* out, size_t n, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH(addConst, n, s, in, c, out, n);
}
// /root/reference/src/filters/AddConstToVectorLength.cpp:97-103
NAIVE_EXPORT cudaError_t gsdrAddToMagnitude(const cuComplex* in, float c, cuComplex* out, size_t n, int32_t cudaDevice, cudaStream_t s) {
  LAUNCH(addToMagnitude, n, s, in, c, out, n);
}
// /root/reference/src/applications/fm_simpletest.cpp:400-413
NAIVE_EXPORT cudaError_t gsdrFmDemod(float rfSampleRate, float tunedFrequency, float channelFrequency, float channelFmDeviation,
                                     size_t D, size_t firstSampleOffset, const float* taps, size_t T, const cuComplex* in,
                                     float* out, size_t nOut, int32_t cudaDevice, cudaStream_t s) {
  const float step = 6.283185307179586f * (tunedFrequency - channelFrequency) / rfSampleRate;
  const float gain = (rfSampleRate / static_cast<float>(D)) / (6.283185307179586f * channelFmDeviation);
  LAUNCH(fmDemod, nOut, s, step, D, firstSampleOffset, taps, T, in, out, gain, nOut);
}
