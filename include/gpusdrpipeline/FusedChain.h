/*
 * FusedChain.h -- ADDITIVE entry point of this library (not in the reference): the whole hot path as ONE graph node.
 *
 *   int8 IQ or complex float  ->  [mix by exp(+j*2*pi*frequency/sampleRate*n)]  ->  decimating FIR (real taps)
 *                             ->  AM / FM quadrature demodulation  ->  audio FIR (real taps, decimating)  ->  float PCM
 *
 * The returned object obeys the reference's Filter contract (filters/Filter.h:41-138): requestBuffer / commitBuffer on
 * port 0 with device memory, readOutput appends float samples.  It replaces the five-node graph the reference builds in
 * src/filters/factories/RfToPcmAudioFactory.cpp:214-304 (plus src/filters/Int8ToFloat.cpp for int8 input) and the
 * hand-driven sequence of src/applications/nbfm_test.cpp:256-354; sample counts, consumption and carried state are
 * those of that cascade (Fir.cpp:141-187, QuadFmDemod.cpp:76-113).
 *
 * IRfToPcmAudioFactory::createRfToPcm() of this library returns the same node with taps it designs itself; this
 * function takes the taps from the caller (identical taps on both sides is what parity tests need).
 */
#ifndef GPUSDRPIPELINE_FUSEDCHAIN_H
#define GPUSDRPIPELINE_FUSEDCHAIN_H

#include <gpusdrpipeline/Factories.h>

struct GsFusedChainParams {
  uint32_t structSize;      /* sizeof(GsFusedChainParams) */
  SampleType inputType;     /* SampleType_Int8Complex (interleaved int8 I,Q) or SampleType_FloatComplex */
  Modulation modulation;    /* Modulation_Am / Modulation_Fm */
  uint32_t mix;             /* 0: no mixer */
  double sampleRate;        /* Hz, of the input */
  double frequency;         /* Hz; the cosine-source frequency (tuned - channel) */
  const float* rfTaps;      /* host pointer, correlation order, as passed to IFirFactory::createFir */
  size_t rfTapCount;
  size_t rfDecimation;
  float fmGain;             /* ignored unless FM; factories/QuadDemodFactory.h:108-110 */
  uint32_t reserved;
  const float* audioTaps;   /* host pointer */
  size_t audioTapCount;
  size_t audioDecimation;
};

GS_EXPORT [[nodiscard]] Result<Filter> gsCreateFusedChain(const GsFusedChainParams* params, ICudaCommandQueue* commandQueue) noexcept;

/* The taps IRfToPcmAudioFactory::createRfToPcm() of this library designs for these parameters (Kaiser-windowed sinc of the
 * reference's own length estimate, RfToPcmAudioFactory.cpp:44-47,164-170; the reference's designer, kernrj/remez-exchange,
 * is not vendored), so that a caller -- and the parity tests -- can run any other implementation with IDENTICAL taps.
 * Either taps pointer may be NULL to query the counts only.  Status_OutOfRange if a capacity is too small. */
GS_EXPORT [[nodiscard]] Status gsDesignRfToPcmTaps(
    float rfSampleRate, size_t rfLowPassDecimation, size_t audioLowPassDecimation, float rfLowPassDbAttenuation,
    float audioLowPassDbAttenuation, float* rfTaps, size_t rfTapCapacity, size_t* rfTapCount, float* audioTaps, size_t audioTapCapacity,
    size_t* audioTapCount) noexcept;

#endif  // GPUSDRPIPELINE_FUSEDCHAIN_H
