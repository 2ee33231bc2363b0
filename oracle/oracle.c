/*
 * oracle.c -- CPU restatement (fp64 arithmetic) of the streaming DSP hot path of
 * kernrj/cuda-sdr (gpusdrpipeline).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  Nothing under cuda_sdr_b200/ links,
 * imports or executes it; the product path fails loudly when the CUDA library is missing.
 *
 * Pinning status.  The arithmetic of the reference lives in the un-vendored, un-pinned third-party
 * library `gsdr` (github.com/kernrj/gsdr @ main, /root/reference/src/CMakeLists.txt:13-19) whose source
 * is absent.  The oracle is therefore pinned as follows:
 *   - FIR (float taps x complex data, decimating), sample counts, consumption and cross-commit history:
 *     PINNED by /root/reference/tests/FirTests.cpp:11-14,39-47,81-84 and :97-101,125-134,196-202
 *     (tests/golden/fir_kat.json, tests/test_oracle_kat.py).
 *   - Complex cosine source: PINNED by /root/reference/tests/CosineSourceTests.cpp:12-13,49-55
 *     (tests/golden/cosine_kat.json).
 *   - int8->float scale, MultiplyCC, QuadAmDemod, QuadFmDemod, FIR FF/CC/CF, real cosine, gsdrFmDemod:
 *     PARITY UNPINNED -- the reference holds no test or golden vector for them; the semantics are
 *     restated from the call sites and names (file:line cited at each function).
 *
 * Conventions: complex arrays are interleaved (re, im).  Inputs are the fp32 / int8 values the GPU
 * sees; all arithmetic and all outputs are fp64 unless the function name says otherwise.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

static const double ORC_TWO_PI = 6.283185307179586476925286766559;

/* ------------------------------------------------------------------------------------------------
 * Host threads actually used by the OpenMP loops (reported as cpu_baseline.cores).
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Sample-count rules (integer, must match exactly).
 * ---------------------------------------------------------------------------------------------- */

/* Fir::getNumOutputElements, /root/reference/src/filters/Fir.cpp:141-187, restated VERBATIM in size_t
 * arithmetic, including its wrap-around when decimation >= tapCount + 2 or nIn < tapCount - 1.
 * Used by tests to show where the well-defined rule below equals the reference. */
ORC_API size_t orc_fir_num_outputs_verbatim(size_t nIn, size_t tapCount, size_t decimation) {
  const size_t need = tapCount - decimation + 1; /* Fir.cpp:181 */
  if (nIn < need) return 0;                     /* Fir.cpp:182-184 */
  return (nIn - (tapCount - 1)) / decimation;   /* Fir.cpp:186 */
}

/* Well-defined count used by the product: equals the verbatim rule wherever that rule does not wrap
 * (SURVEY.md section 8(a) row a4). */
ORC_API size_t orc_fir_num_outputs(size_t nIn, size_t tapCount, size_t decimation) {
  if (decimation == 0) decimation = 1; /* Fir.cpp:119 */
  if (tapCount == 0) return 0;
  if (nIn + 1 < tapCount) return 0;
  return (nIn + 1 - tapCount) / decimation;
}

/* QuadFmDemod keeps one sample of history: n outputs need n+1 inputs
 * (/root/reference/src/filters/QuadFmDemod.cpp:76-84,92-96). */
ORC_API size_t orc_fm_num_outputs(size_t nIn) { return nIn == 0 ? 0 : nIn - 1; }

/* FM gain of the factory, /root/reference/src/filters/factories/QuadDemodFactory.h:108-110
 * (computed in float exactly as written there). */
ORC_API float orc_fm_gain(float inputSampleRate, float fskDeviation) {
  return inputSampleRate / (2.0f * (float)M_PI * fskDeviation * 5);
}

/* ------------------------------------------------------------------------------------------------
 * Element-wise ops.
 * ---------------------------------------------------------------------------------------------- */

/* gsdrInt8ToNormFloat, called at /root/reference/src/filters/Int8ToFloat.cpp:89-94.  Scale INFERRED
 * (1/128: exact in fp32; PARITY UNPINNED).  Output is fp32 because the gate is bit-exactness. */
ORC_API void orc_int8_to_norm_float(const int8_t* in, float* out, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) out[i] = (float)in[i] * (1.0f / 128.0f);
}

/* gsdrCosineC, called at /root/reference/src/filters/ComplexCosineSource.cpp:74-80 with
 * phiEnd = mPhi + n*delta (:72).  out[i] = exp(j*(phiStart + i*(phiEnd-phiStart)/n)).
 * PINNED by tests/CosineSourceTests.cpp:49-55 to 1e-4. */
ORC_API void orc_cosine_c(float phiStart, float phiEnd, double* out, size_t n) {
  const double p0 = (double)phiStart;
  const double step = n ? ((double)phiEnd - (double)phiStart) / (double)n : 0.0;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    const double phi = p0 + (double)i * step;
    out[2 * i] = cos(phi);
    out[2 * i + 1] = sin(phi);
  }
}

/* gsdrCosineF, called at /root/reference/src/filters/CosineSource.cpp:74-80. PARITY UNPINNED. */
ORC_API void orc_cosine_f(float phiStart, float phiEnd, double* out, size_t n) {
  const double p0 = (double)phiStart;
  const double step = n ? ((double)phiEnd - (double)phiStart) / (double)n : 0.0;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) out[i] = cos(p0 + (double)i * step);
}

/* Host-side phase bookkeeping of the cosine sources: CosineSource.cpp:51,72,82 /
 * ComplexCosineSource.cpp:52,72,82 -- float32 end to end, fmod by 2*pi_f. */
ORC_API float orc_cosine_delta(float sampleRate, float frequency) {
  return (float)(2.0 * M_PI * frequency / sampleRate); /* ComplexCosineSource.cpp:52 */
}
ORC_API float orc_cosine_phi_end(float phi, size_t n, float delta) {
  return phi + (float)n * delta; /* ComplexCosineSource.cpp:72 */
}
ORC_API float orc_cosine_next_phi(float phiEnd) {
  return fmodf(phiEnd, 2.0f * (float)M_PI); /* ComplexCosineSource.cpp:82 */
}

/* gsdrMultiplyCC, called at /root/reference/src/filters/Multiply.cpp:145-151. PARITY UNPINNED. */
ORC_API void orc_multiply_cc(const float* a, const float* b, double* out, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    const double ar = a[2 * i], ai = a[2 * i + 1], br = b[2 * i], bi = b[2 * i + 1];
    out[2 * i] = ar * br - ai * bi;
    out[2 * i + 1] = ar * bi + ai * br;
  }
}

/* gsdrQuadAmDemod / gsdrMagnitude, called at QuadAmDemod.cpp:93-98 / Magnitude.cpp:91-96: |z|.
 * PARITY UNPINNED. */
ORC_API void orc_quad_am_demod(const float* in, double* out, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) out[i] = hypot((double)in[2 * i], (double)in[2 * i + 1]);
}

/* gsdrQuadFmDemod, called at /root/reference/src/filters/QuadFmDemod.cpp:98-104:
 * out[i] = gain * arg(in[i+1] * conj(in[i])), reads n+1 inputs. PARITY UNPINNED. */
ORC_API void orc_quad_fm_demod(const float* in, double* out, float gain, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    const double ar = in[2 * i], ai = in[2 * i + 1], br = in[2 * i + 2], bi = in[2 * i + 3];
    const double re = br * ar + bi * ai;
    const double im = bi * ar - br * ai;
    out[i] = (double)gain * atan2(im, re);
  }
}

/* gsdrAddConstFF, called at /root/reference/src/filters/AddConst.cpp:99-105. PARITY UNPINNED. */
ORC_API void orc_add_const_ff(const float* in, float c, double* out, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) out[i] = (double)in[i] + (double)c;
}

/* gsdrAddToMagnitude, called at /root/reference/src/filters/AddConstToVectorLength.cpp:97-103:
 * z * (|z| + c) / |z|.  PARITY UNPINNED (|z| == 0 -> 0 here). */
ORC_API void orc_add_to_magnitude(const float* in, float c, double* out, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    const double re = in[2 * i], im = in[2 * i + 1];
    const double mag = hypot(re, im);
    const double s = mag > 0.0 ? (mag + (double)c) / mag : 0.0;
    out[2 * i] = re * s;
    out[2 * i + 1] = im * s;
  }
}

/* ------------------------------------------------------------------------------------------------
 * FIR: y[k] = sum_{j<T} h[j] * x[k*D + j]  (correlation order, taps used as given -- the caller
 * pre-reverses: /root/reference/src/filters/Fir.cpp:124 `tapsReversed`).  Pinned by FirTests.cpp.
 * ---------------------------------------------------------------------------------------------- */

/* gsdrFirFC, called at Fir.cpp:240-248: float taps, complex data. */
ORC_API void orc_fir_fc(size_t D, const float* taps, size_t T, const float* in, double* out, size_t nOut) {
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < nOut; k++) {
    const float* x = in + 2 * k * D;
    double re = 0.0, im = 0.0;
    for (size_t j = 0; j < T; j++) {
      const double h = taps[j];
      re += h * (double)x[2 * j];
      im += h * (double)x[2 * j + 1];
    }
    out[2 * k] = re;
    out[2 * k + 1] = im;
  }
}

/* gsdrFirFF, called at Fir.cpp:230-238: float taps, float data. PARITY UNPINNED (same formula). */
ORC_API void orc_fir_ff(size_t D, const float* taps, size_t T, const float* in, double* out, size_t nOut) {
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < nOut; k++) {
    const float* x = in + k * D;
    double acc = 0.0;
    for (size_t j = 0; j < T; j++) acc += (double)taps[j] * (double)x[j];
    out[k] = acc;
  }
}

/* gsdrFirCC, called at Fir.cpp:250-258: complex taps, complex data. PARITY UNPINNED. */
ORC_API void orc_fir_cc(size_t D, const float* taps, size_t T, const float* in, double* out, size_t nOut) {
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < nOut; k++) {
    const float* x = in + 2 * k * D;
    double re = 0.0, im = 0.0;
    for (size_t j = 0; j < T; j++) {
      const double hr = taps[2 * j], hi = taps[2 * j + 1];
      const double xr = x[2 * j], xi = x[2 * j + 1];
      re += hr * xr - hi * xi;
      im += hr * xi + hi * xr;
    }
    out[2 * k] = re;
    out[2 * k + 1] = im;
  }
}

/* gsdrFirCF, called at Fir.cpp:260-268: complex taps, float data, complex out. PARITY UNPINNED. */
ORC_API void orc_fir_cf(size_t D, const float* taps, size_t T, const float* in, double* out, size_t nOut) {
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < nOut; k++) {
    const float* x = in + k * D;
    double re = 0.0, im = 0.0;
    for (size_t j = 0; j < T; j++) {
      re += (double)taps[2 * j] * (double)x[j];
      im += (double)taps[2 * j + 1] * (double)x[j];
    }
    out[2 * k] = re;
    out[2 * k + 1] = im;
  }
}

/* Same as the four above but with fp64 input (used inside the chain oracle, where the upstream
 * stage's fp64 result feeds the next stage without an fp32 rounding). */
static void fir_real_taps_c64(size_t D, const float* taps, size_t T, const double* in, double* out, size_t nOut) {
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < nOut; k++) {
    const double* x = in + 2 * k * D;
    double re = 0.0, im = 0.0;
    for (size_t j = 0; j < T; j++) {
      const double h = taps[j];
      re += h * x[2 * j];
      im += h * x[2 * j + 1];
    }
    out[2 * k] = re;
    out[2 * k + 1] = im;
  }
}

static void fir_real_taps_r64(size_t D, const float* taps, size_t T, const double* in, double* out, size_t nOut) {
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < nOut; k++) {
    const double* x = in + k * D;
    double acc = 0.0;
    for (size_t j = 0; j < T; j++) acc += (double)taps[j] * x[j];
    out[k] = acc;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Exact mixer phase: phase of absolute sample n is 2*pi*frac(n * cyclesPerSample), evaluated with a
 * 64-bit fixed-point turn accumulator so that it does not degrade with n (SURVEY.md section 0 fact 5:
 * the reference's own float32 phase does; see DESIGN.md "phase modes").  `phaseStep` is
 * round(frac(f/fs) * 2^64) as an unsigned 64-bit integer, two's complement for negative f.
 * ---------------------------------------------------------------------------------------------- */
ORC_API uint64_t orc_phase_step(double frequency, double sampleRate) {
  long double cyc = (long double)frequency / (long double)sampleRate;
  cyc -= floorl(cyc); /* [0,1) */
  long double scaled = cyc * 18446744073709551616.0L;
  if (scaled >= 18446744073709551615.0L) return 0;
  return (uint64_t)llroundl(scaled >= 9223372036854775808.0L ? scaled - 18446744073709551616.0L : scaled);
}

static inline void exact_phasor(uint64_t phaseStep, uint64_t n, double* c, double* s) {
  const uint64_t turns = phaseStep * n; /* mod 2^64 */
  /* signed interpretation keeps the argument in [-pi, pi) for best sincos accuracy */
  const double frac = (double)(int64_t)turns * (1.0 / 18446744073709551616.0);
  const double phi = ORC_TWO_PI * frac;
  *c = cos(phi);
  *s = sin(phi);
}

/* ------------------------------------------------------------------------------------------------
 * The fused chain, staged exactly like the reference graph
 *   Int8ToFloat -> MultiplyCcc(ComplexCosineSource) -> Fir FC (D1) -> QuadAm|QuadFm -> Fir FF (D2)
 * (/root/reference/src/applications/nbfm_test.cpp:256-354, RfToPcmAudioFactory.cpp:214-304), all in
 * fp64, over one contiguous block whose first sample has absolute index n0.
 *
 *   modulation: 0 = AM (|y|), 1 = FM (gain*arg(y[k+1] conj y[k])), 2 = none (stop after RF FIR; the
 *               complex RF FIR output is returned in rfOut and audioOut is untouched).
 *   taps2 == NULL or T2 == 0: no audio FIR; audioOut receives the demodulated samples.
 *   inputIsInt8: 1 -> `in` is int8 IQ pairs, scaled by 1/128; 0 -> `in` is float32 IQ pairs.
 *   phaseStep == 0 and mix == 0: mixer skipped.
 * Returns the number of audio (final) outputs written; counts follow orc_fir_num_outputs /
 * orc_fm_num_outputs.  rfOut (optional, may be NULL) receives the complex RF FIR outputs, demodOut
 * (optional) the demodulated samples.
 * ---------------------------------------------------------------------------------------------- */
ORC_API size_t orc_chain(
    const void* in,
    int inputIsInt8,
    size_t n,
    uint64_t n0,
    int mix,
    uint64_t phaseStep,
    const float* taps1,
    size_t T1,
    size_t D1,
    int modulation,
    float fmGain,
    const float* taps2,
    size_t T2,
    size_t D2,
    double* rfOut,
    double* demodOut,
    double* audioOut) {
  const size_t nRf = orc_fir_num_outputs(n, T1, D1);
  if (nRf == 0) return 0;
  if (D1 == 0) D1 = 1;

  /* only the samples the RF FIR actually reads */
  const size_t nUsed = (nRf - 1) * D1 + T1;
  double* z = (double*)malloc(sizeof(double) * 2 * nUsed);
  if (!z) return 0;

  const int8_t* in8 = (const int8_t*)in;
  const float* inF = (const float*)in;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nUsed; i++) {
    double xr, xi;
    if (inputIsInt8) {
      xr = (double)in8[2 * i] * (1.0 / 128.0);
      xi = (double)in8[2 * i + 1] * (1.0 / 128.0);
    } else {
      xr = inF[2 * i];
      xi = inF[2 * i + 1];
    }
    if (mix) {
      double c, s;
      exact_phasor(phaseStep, n0 + (uint64_t)i, &c, &s);
      z[2 * i] = xr * c - xi * s;
      z[2 * i + 1] = xr * s + xi * c;
    } else {
      z[2 * i] = xr;
      z[2 * i + 1] = xi;
    }
  }

  double* y = rfOut ? rfOut : (double*)malloc(sizeof(double) * 2 * nRf);
  if (!y) {
    free(z);
    return 0;
  }
  fir_real_taps_c64(D1, taps1, T1, z, y, nRf);
  free(z);

  if (modulation == 2) {
    if (!rfOut) free(y);
    return nRf;
  }

  const size_t nDemod = modulation == 1 ? orc_fm_num_outputs(nRf) : nRf;
  const int haveAudioFir = taps2 != NULL && T2 > 0;
  double* d = demodOut ? demodOut : (haveAudioFir ? (double*)malloc(sizeof(double) * (nDemod ? nDemod : 1)) : audioOut);
  if (!d) {
    if (!rfOut) free(y);
    return 0;
  }

#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < nDemod; k++) {
    if (modulation == 0) {
      d[k] = hypot(y[2 * k], y[2 * k + 1]);
    } else {
      const double ar = y[2 * k], ai = y[2 * k + 1], br = y[2 * k + 2], bi = y[2 * k + 3];
      d[k] = (double)fmGain * atan2(bi * ar - br * ai, br * ar + bi * ai);
    }
  }
  if (!rfOut) free(y);

  size_t nAudio;
  if (haveAudioFir) {
    nAudio = orc_fir_num_outputs(nDemod, T2, D2);
    if (D2 == 0) D2 = 1;
    fir_real_taps_r64(D2, taps2, T2, d, audioOut, nAudio);
    if (!demodOut) free(d);
  } else {
    nAudio = nDemod;
    if (demodOut && audioOut) memcpy(audioOut, d, sizeof(double) * nDemod);
  }
  return nAudio;
}

/* Number of final outputs orc_chain() produces for n input samples (same rules, no arithmetic). */
ORC_API size_t orc_chain_num_outputs(size_t n, size_t T1, size_t D1, int modulation, size_t T2, size_t D2) {
  const size_t nRf = orc_fir_num_outputs(n, T1, D1);
  if (modulation == 2) return nRf;
  const size_t nDemod = modulation == 1 ? orc_fm_num_outputs(nRf) : nRf;
  if (T2 == 0) return nDemod;
  return orc_fir_num_outputs(nDemod, T2, D2);
}

/* gsdrFmDemod (the upstream author's own fused kernel), called at
 * /root/reference/src/applications/fm_simpletest.cpp:400-413: mix by (tuned - channel), real-tap
 * low-pass, decimate, quadrature-FM demod in one call.  PARITY UNPINNED; restated as the chain above
 * with the mixer phase index starting at firstSampleOffset and gain = fs_out/(2*pi*deviation). */
ORC_API size_t orc_fm_demod_fused(
    float rfSampleRate,
    float tunedFrequency,
    float channelFrequency,
    float channelFmDeviation,
    size_t decimation,
    size_t firstSampleOffset,
    const float* taps,
    size_t tapCount,
    const float* in,
    size_t nIn,
    double* out) {
  const uint64_t step = orc_phase_step((double)tunedFrequency - (double)channelFrequency, (double)rfSampleRate);
  const float gain = (rfSampleRate / (float)decimation) / (2.0f * (float)M_PI * channelFmDeviation);
  return orc_chain(in, 0, nIn, (uint64_t)firstSampleOffset, 1, step, taps, tapCount, decimation, 1, gain, NULL, 0, 1, NULL, NULL, out);
}
