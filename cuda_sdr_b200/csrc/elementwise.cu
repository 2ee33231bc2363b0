// Element-wise kernels of the gsdr C-ABI (include/gsdr/gsdr.h, include/gsdr/conversion.h).
// All are HBM-bound streaming kernels: 128-bit coalesced accesses, L1 no-allocate, grid-stride loops
// over a grid that is a multiple of the SM count.
#include <gsdr/conversion.h>
#include <gsdr/gsdr.h>

#include "common.cuh"

namespace b200sdr {

std::atomic<uint64_t> g_launchCount {0};

namespace {

constexpr unsigned kThreads = 256;

// ---- int8 -> float (scale 1/128) -----------------------------------------------------------------
// Algorithmic bytes: 1 read + 4 written per scalar.
__global__ void __launch_bounds__(kThreads) int8ToNormFloatVec(const uint4* __restrict__ in, float4* __restrict__ out, size_t groups) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const uint4 v = ldStream(in + g);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float4 f;
      int8x4ToFloat(w[k], f.x, f.y, f.z, f.w);
      f.x *= 0.0078125f;
      f.y *= 0.0078125f;
      f.z *= 0.0078125f;
      f.w *= 0.0078125f;
      stStream(out + g * 4 + k, f);
    }
  }
}

__global__ void __launch_bounds__(kThreads) int8ToNormFloatScalar(const int8_t* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    out[i] = static_cast<float>(in[i]) * 0.0078125f;
  }
}

// ---- cosine sources --------------------------------------------------------------------------------
// phi_i = phiStart + i * (phiEnd - phiStart) / n, evaluated in fp64 and reduced to [-pi, pi] before the
// fp32 sincos so that the kernel adds no phase error of its own to the (float32) phase the caller hands in.
__device__ __forceinline__ float reducedPhase(double phi0, double step, size_t i) {
  constexpr double kTwoPi = 6.283185307179586476925286766559;
  constexpr double kInvTwoPi = 0.15915494309189533576888376337251;
  const double phi = fma(static_cast<double>(i), step, phi0);
  return static_cast<float>(fma(-rint(phi * kInvTwoPi), kTwoPi, phi));
}

__global__ void __launch_bounds__(kThreads) cosineComplex(double phi0, double step, float2* __restrict__ out, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float s, c;
    sincosf(reducedPhase(phi0, step, i), &s, &c);
    out[i] = make_float2(c, s);
  }
}

__global__ void __launch_bounds__(kThreads) cosineReal(double phi0, double step, float* __restrict__ out, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    out[i] = cosf(reducedPhase(phi0, step, i));
  }
}

// ---- complex multiply ------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y));
}

__global__ void __launch_bounds__(kThreads) multiplyComplexVec(
    const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out, size_t pairs) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < pairs; i += stride) {
    const float4 x = ldStream(a + i);
    const float4 y = ldStream(b + i);
    const float2 p = cmul(make_float2(x.x, x.y), make_float2(y.x, y.y));
    const float2 q = cmul(make_float2(x.z, x.w), make_float2(y.z, y.w));
    stStream(out + i, make_float4(p.x, p.y, q.x, q.y));
  }
}

__global__ void __launch_bounds__(kThreads) multiplyComplexScalar(
    const float2* __restrict__ a, const float2* __restrict__ b, float2* __restrict__ out, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    out[i] = cmul(a[i], b[i]);
  }
}

// ---- magnitude / AM demod --------------------------------------------------------------------------
__device__ __forceinline__ float mag(float re, float im) { return sqrtf(fmaf(re, re, im * im)); }

__global__ void __launch_bounds__(kThreads) magnitudeVec(const float4* __restrict__ in, float4* __restrict__ out, size_t quads) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < quads; i += stride) {
    const float4 x = ldStream(in + 2 * i);
    const float4 y = ldStream(in + 2 * i + 1);
    stStream(out + i, make_float4(mag(x.x, x.y), mag(x.z, x.w), mag(y.x, y.y), mag(y.z, y.w)));
  }
}

__global__ void __launch_bounds__(kThreads) magnitudeScalar(const float2* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float2 z = in[i];
    out[i] = mag(z.x, z.y);
  }
}

// ---- quadrature FM demod ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) quadFmDemod(const float2* __restrict__ in, float* __restrict__ out, float gain, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float2 a = in[i];
    const float2 b = in[i + 1];
    const float re = fmaf(b.y, a.y, b.x * a.x);
    const float im = fmaf(b.y, a.x, -b.x * a.y);
    out[i] = gain * atan2f(im, re);
  }
}

// ---- add const / add to magnitude ------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) addConstScalar(const float* __restrict__ in, float c, float* __restrict__ out, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i] + c;
}

__global__ void __launch_bounds__(kThreads) addConstVec(const float4* __restrict__ in, float c, float4* __restrict__ out, size_t quads) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < quads; i += stride) {
    const float4 x = ldStream(in + i);
    stStream(out + i, make_float4(x.x + c, x.y + c, x.z + c, x.w + c));
  }
}

__global__ void __launch_bounds__(kThreads) addToMagnitude(const float2* __restrict__ in, float c, float2* __restrict__ out, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float2 z = in[i];
    const float m = mag(z.x, z.y);
    const float s = m > 0.0f ? (m + c) / m : 0.0f;
    out[i] = make_float2(z.x * s, z.y * s);
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace b200sdr

using namespace b200sdr;

#define GSDR_ENTER(device)                                   \
  DeviceGuard guard__(device);                               \
  if (guard__.status != cudaSuccess) return guard__.status;  \
  if (numElements == 0) return cudaSuccess

GSDR_EXPORT cudaError_t gsdrInt8ToNormFloat(
    const int8_t* input, float* output, size_t numElements, int32_t cudaDevice, cudaStream_t cudaStream) {
  GSDR_ENTER(cudaDevice);
  size_t done = 0;
  if (aligned16(input) && aligned16(output) && numElements >= 16) {
    const size_t groups = numElements / 16;
    int8ToNormFloatVec<<<elementwiseGrid(groups, kThreads), kThreads, 0, cudaStream>>>(
        reinterpret_cast<const uint4*>(input), reinterpret_cast<float4*>(output), groups);
    const cudaError_t e = launchStatus();
    if (e != cudaSuccess) return e;
    done = groups * 16;
  }
  if (done < numElements) {
    const size_t rest = numElements - done;
    int8ToNormFloatScalar<<<elementwiseGrid(rest, kThreads), kThreads, 0, cudaStream>>>(input + done, output + done, rest);
    return launchStatus();
  }
  return cudaSuccess;
}

GSDR_EXPORT cudaError_t gsdrCosineC(
    float phiStart, float phiEnd, cuComplex* output, size_t numElements, int32_t cudaDevice, cudaStream_t cudaStream) {
  GSDR_ENTER(cudaDevice);
  const double step = (static_cast<double>(phiEnd) - static_cast<double>(phiStart)) / static_cast<double>(numElements);
  cosineComplex<<<elementwiseGrid(numElements, kThreads), kThreads, 0, cudaStream>>>(
      static_cast<double>(phiStart), step, reinterpret_cast<float2*>(output), numElements);
  return launchStatus();
}

GSDR_EXPORT cudaError_t gsdrCosineF(
    float phiStart, float phiEnd, float* output, size_t numElements, int32_t cudaDevice, cudaStream_t cudaStream) {
  GSDR_ENTER(cudaDevice);
  const double step = (static_cast<double>(phiEnd) - static_cast<double>(phiStart)) / static_cast<double>(numElements);
  cosineReal<<<elementwiseGrid(numElements, kThreads), kThreads, 0, cudaStream>>>(
      static_cast<double>(phiStart), step, output, numElements);
  return launchStatus();
}

GSDR_EXPORT cudaError_t gsdrMultiplyCC(
    const cuComplex* a, const cuComplex* b, cuComplex* output, size_t numElements, int32_t cudaDevice,
    cudaStream_t cudaStream) {
  GSDR_ENTER(cudaDevice);
  size_t done = 0;
  if (aligned16(a) && aligned16(b) && aligned16(output) && numElements >= 2) {
    const size_t pairs = numElements / 2;
    multiplyComplexVec<<<elementwiseGrid(pairs, kThreads), kThreads, 0, cudaStream>>>(
        reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), reinterpret_cast<float4*>(output), pairs);
    const cudaError_t e = launchStatus();
    if (e != cudaSuccess) return e;
    done = pairs * 2;
  }
  if (done < numElements) {
    const size_t rest = numElements - done;
    multiplyComplexScalar<<<elementwiseGrid(rest, kThreads), kThreads, 0, cudaStream>>>(
        reinterpret_cast<const float2*>(a) + done, reinterpret_cast<const float2*>(b) + done,
        reinterpret_cast<float2*>(output) + done, rest);
    return launchStatus();
  }
  return cudaSuccess;
}

static cudaError_t magnitudeImpl(const cuComplex* input, float* output, size_t numElements, cudaStream_t cudaStream) {
  size_t done = 0;
  if (aligned16(input) && aligned16(output) && numElements >= 4) {
    const size_t quads = numElements / 4;
    magnitudeVec<<<elementwiseGrid(quads, kThreads), kThreads, 0, cudaStream>>>(
        reinterpret_cast<const float4*>(input), reinterpret_cast<float4*>(output), quads);
    const cudaError_t e = launchStatus();
    if (e != cudaSuccess) return e;
    done = quads * 4;
  }
  if (done < numElements) {
    const size_t rest = numElements - done;
    magnitudeScalar<<<elementwiseGrid(rest, kThreads), kThreads, 0, cudaStream>>>(
        reinterpret_cast<const float2*>(input) + done, output + done, rest);
    return launchStatus();
  }
  return cudaSuccess;
}

GSDR_EXPORT cudaError_t gsdrQuadAmDemod(
    const cuComplex* input, float* output, size_t numElements, int32_t cudaDevice, cudaStream_t cudaStream) {
  GSDR_ENTER(cudaDevice);
  return magnitudeImpl(input, output, numElements, cudaStream);
}

GSDR_EXPORT cudaError_t gsdrMagnitude(
    const cuComplex* input, float* output, size_t numElements, int32_t cudaDevice, cudaStream_t cudaStream) {
  GSDR_ENTER(cudaDevice);
  return magnitudeImpl(input, output, numElements, cudaStream);
}

GSDR_EXPORT cudaError_t gsdrQuadFmDemod(
    const cuComplex* input, float* output, float gain, size_t numElements, int32_t cudaDevice,
    cudaStream_t cudaStream) {
  GSDR_ENTER(cudaDevice);
  quadFmDemod<<<elementwiseGrid(numElements, kThreads), kThreads, 0, cudaStream>>>(
      reinterpret_cast<const float2*>(input), output, gain, numElements);
  return launchStatus();
}

GSDR_EXPORT cudaError_t gsdrAddConstFF(
    const float* input, float addConst, float* output, size_t numElements, int32_t cudaDevice,
    cudaStream_t cudaStream) {
  GSDR_ENTER(cudaDevice);
  size_t done = 0;
  if (aligned16(input) && aligned16(output) && numElements >= 4) {
    const size_t quads = numElements / 4;
    addConstVec<<<elementwiseGrid(quads, kThreads), kThreads, 0, cudaStream>>>(
        reinterpret_cast<const float4*>(input), addConst, reinterpret_cast<float4*>(output), quads);
    const cudaError_t e = launchStatus();
    if (e != cudaSuccess) return e;
    done = quads * 4;
  }
  if (done < numElements) {
    const size_t rest = numElements - done;
    addConstScalar<<<elementwiseGrid(rest, kThreads), kThreads, 0, cudaStream>>>(input + done, addConst, output + done, rest);
    return launchStatus();
  }
  return cudaSuccess;
}

GSDR_EXPORT cudaError_t gsdrAddToMagnitude(
    const cuComplex* input, float addToMagnitudeValue, cuComplex* output, size_t numElements, int32_t cudaDevice,
    cudaStream_t cudaStream) {
  GSDR_ENTER(cudaDevice);
  addToMagnitude<<<elementwiseGrid(numElements, kThreads), kThreads, 0, cudaStream>>>(
      reinterpret_cast<const float2*>(input), addToMagnitudeValue, reinterpret_cast<float2*>(output), numElements);
  return launchStatus();
}
