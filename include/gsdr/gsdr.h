/*
 * gsdr.h -- C-ABI of the device kernels behind the streaming DSP hot path.
 *
 * Drop-in for the `gsdr` library the reference's filter wrappers bind (kernrj/cuda-sdr,
 * src/CMakeLists.txt:13-19: un-vendored, un-pinned; source absent).  Every entry point below has
 * exactly the name, argument order and return type the reference call site uses, so the reference's
 * own src/filters/*.cpp compile and link against this library unchanged.  All pointers are DEVICE
 * pointers; work is enqueued on `cudaStream` of device `cudaDevice`; nothing synchronises.
 *
 * Implemented from scratch as sm_100a kernels in cuda_sdr_b200/csrc/.  No CPU fallback: without a
 * CUDA device every function returns the CUDA error of the failed launch.
 */
#ifndef B200SDR_GSDR_GSDR_H
#define B200SDR_GSDR_GSDR_H

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define GSDR_C_LINKAGE extern "C"
#else
#define GSDR_C_LINKAGE
#endif
#define GSDR_EXPORT GSDR_C_LINKAGE __attribute__((visibility("default")))

/* out[i] = cos(phiStart + i*(phiEnd-phiStart)/n).   Replaces the call at
 * reference src/filters/CosineSource.cpp:74-80. */
GSDR_EXPORT cudaError_t gsdrCosineF(
    float phiStart, float phiEnd, float* output, size_t numElements, int32_t cudaDevice, cudaStream_t cudaStream);

/* out[i] = (cos phi_i, sin phi_i).   Replaces src/filters/ComplexCosineSource.cpp:74-80
 * (pinned by tests/CosineSourceTests.cpp:49-55). */
GSDR_EXPORT cudaError_t gsdrCosineC(
    float phiStart, float phiEnd, cuComplex* output, size_t numElements, int32_t cudaDevice, cudaStream_t cudaStream);

/* out[i] = a[i] * b[i] (complex).   Replaces src/filters/Multiply.cpp:145-151. */
GSDR_EXPORT cudaError_t gsdrMultiplyCC(
    const cuComplex* a, const cuComplex* b, cuComplex* output, size_t numElements, int32_t cudaDevice,
    cudaStream_t cudaStream);

/* Decimating FIR, correlation order, taps used as given:
 *   out[k] = sum_{j < tapCount} taps[j] * in[k*decimation + j],  k < numOutputs.
 * Reads (numOutputs-1)*decimation + tapCount input elements.
 * FF: float taps, float data   -- replaces src/filters/Fir.cpp:230-238
 * FC: float taps, complex data -- replaces src/filters/Fir.cpp:240-248 (pinned by tests/FirTests.cpp)
 * CC: complex taps, complex data -- replaces src/filters/Fir.cpp:250-258
 * CF: complex taps, float data, complex out -- replaces src/filters/Fir.cpp:260-268 */
GSDR_EXPORT cudaError_t gsdrFirFF(
    size_t decimation, const float* taps, size_t tapCount, const float* input, float* output, size_t numOutputs,
    int32_t cudaDevice, cudaStream_t cudaStream);
GSDR_EXPORT cudaError_t gsdrFirFC(
    size_t decimation, const float* taps, size_t tapCount, const cuComplex* input, cuComplex* output,
    size_t numOutputs, int32_t cudaDevice, cudaStream_t cudaStream);
GSDR_EXPORT cudaError_t gsdrFirCC(
    size_t decimation, const cuComplex* taps, size_t tapCount, const cuComplex* input, cuComplex* output,
    size_t numOutputs, int32_t cudaDevice, cudaStream_t cudaStream);
GSDR_EXPORT cudaError_t gsdrFirCF(
    size_t decimation, const cuComplex* taps, size_t tapCount, const float* input, cuComplex* output,
    size_t numOutputs, int32_t cudaDevice, cudaStream_t cudaStream);

/* out[i] = |in[i]|.   Replaces src/filters/QuadAmDemod.cpp:93-98. */
GSDR_EXPORT cudaError_t gsdrQuadAmDemod(
    const cuComplex* input, float* output, size_t numElements, int32_t cudaDevice, cudaStream_t cudaStream);

/* out[i] = gain * arg(in[i+1] * conj(in[i])); reads numElements+1 inputs.
 * Replaces src/filters/QuadFmDemod.cpp:98-104. */
GSDR_EXPORT cudaError_t gsdrQuadFmDemod(
    const cuComplex* input, float* output, float gain, size_t numElements, int32_t cudaDevice,
    cudaStream_t cudaStream);

/* out[i] = |in[i]|.   Replaces src/filters/Magnitude.cpp:91-96. */
GSDR_EXPORT cudaError_t gsdrMagnitude(
    const cuComplex* input, float* output, size_t numElements, int32_t cudaDevice, cudaStream_t cudaStream);

/* out[i] = in[i] + addConst.   Replaces src/filters/AddConst.cpp:99-105. */
GSDR_EXPORT cudaError_t gsdrAddConstFF(
    const float* input, float addConst, float* output, size_t numElements, int32_t cudaDevice,
    cudaStream_t cudaStream);

/* out[i] = in[i] * (|in[i]| + addToMagnitude) / |in[i]|.   Replaces
 * src/filters/AddConstToVectorLength.cpp:97-103. */
GSDR_EXPORT cudaError_t gsdrAddToMagnitude(
    const cuComplex* input, float addToMagnitude, cuComplex* output, size_t numElements, int32_t cudaDevice,
    cudaStream_t cudaStream);

/* The upstream author's fused op: mix by (tuned-channel), real-tap low-pass, decimate, quadrature-FM
 * demodulate, in one launch.  Replaces src/applications/fm_simpletest.cpp:400-413.
 * Reads (outputCount)*decimation + tapCount complex inputs (one extra FIR output for the discriminator);
 * mixer phase index of input[0] is firstSampleOffset; gain = (fs/decimation)/(2*pi*deviation). */
GSDR_EXPORT cudaError_t gsdrFmDemod(
    float rfSampleRate, float tunedFrequency, float channelFrequency, float channelFmDeviation,
    size_t rfLowPassDecimation, size_t firstSampleOffset, const float* lowPassTaps, size_t lowPassTapCount,
    const cuComplex* input, float* output, size_t outputCount, int32_t cudaDevice, cudaStream_t cudaStream);

#endif /* B200SDR_GSDR_GSDR_H */
