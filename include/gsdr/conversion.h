/*
 * conversion.h -- sample-format conversion entry points of the gsdr C-ABI (see gsdr.h).
 */
#ifndef B200SDR_GSDR_CONVERSION_H
#define B200SDR_GSDR_CONVERSION_H

#include <gsdr/gsdr.h>

/* out[i] = float(in[i]) / 128 for numElements SCALARS (I and Q count separately); exact in fp32.
 * Replaces the call at reference src/filters/Int8ToFloat.cpp:89-94. */
GSDR_EXPORT cudaError_t gsdrInt8ToNormFloat(
    const int8_t* input, float* output, size_t numElements, int32_t cudaDevice, cudaStream_t cudaStream);

#endif /* B200SDR_GSDR_CONVERSION_H */
