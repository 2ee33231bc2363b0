/*
 * b200sdr.h -- C-ABI of the fused int8/cf32 -> mix -> decimating FIR -> AM/FM demod -> audio FIR chain.
 *
 * Plain pointers and sizes only.  This is the boundary the host framework (libgpusdrpipeline:
 * getFactoriesSingleton()/IFactories, include/gpusdrpipeline/) and any foreign-language binding sit on.
 * One `b200sdr_chain` replaces, as ONE node, the five-node graph the reference builds in
 *   src/filters/factories/RfToPcmAudioFactory.cpp:214-304  (cosine -> multiply -> Fir -> QuadDemod -> Fir)
 * preceded by src/filters/Int8ToFloat.cpp, i.e. the reference call sequence
 *   src/applications/nbfm_test.cpp:256-354.
 *
 * Status codes are the reference's `Status` values (include/gpusdrpipeline/Status.h:22-34).
 * No CPU fallback exists: every compute entry point needs a CUDA device and fails loudly without one.
 */
#ifndef B200SDR_B200SDR_H
#define B200SDR_B200SDR_H

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define B200SDR_C_LINKAGE extern "C"
#else
#define B200SDR_C_LINKAGE
#endif
#define B200SDR_EXPORT B200SDR_C_LINKAGE __attribute__((visibility("default")))

typedef uint32_t b200sdr_status; /* Status.h:22-34 */
enum {
  B200SDR_OK = 0,
  B200SDR_UNKNOWN_ERROR = 1,
  B200SDR_OUT_OF_MEMORY = 2,
  B200SDR_RUNTIME_ERROR = 3,
  B200SDR_INVALID_ARGUMENT = 4,
  B200SDR_INVALID_STATE = 5,
  B200SDR_OUT_OF_RANGE = 6,
};

/* input_type uses the reference's SampleType values (include/gpusdrpipeline/SampleType.h:20-25) */
enum { B200SDR_INPUT_CF32 = 0, B200SDR_INPUT_INT8 = 2 };
/* modulation uses the reference's Modulation values (Modulation.h:22-26) plus "none" */
enum { B200SDR_MOD_AM = 0, B200SDR_MOD_FM = 1, B200SDR_MOD_NONE = 2 };

typedef struct b200sdr_chain_config {
  uint32_t struct_size;  /* sizeof(b200sdr_chain_config), for forward compatibility */
  uint32_t input_type;   /* B200SDR_INPUT_*: interleaved int8 IQ (HackRF format) or cuComplex */
  uint32_t modulation;   /* B200SDR_MOD_* */
  uint32_t mix;          /* 0: no mixer; 1: multiply by exp(+j*2*pi*frequency/sample_rate*n) */
  double sample_rate;    /* Hz, of the input */
  double frequency;      /* Hz; the cosine-source frequency (tuned - channel), RfToPcmAudioFactory.cpp:225 */
  const float* rf_taps;  /* HOST pointer, correlation order (as passed to IFirFactory::createFir) */
  size_t rf_tap_count;
  size_t rf_decimation;
  float fm_gain;         /* QuadDemodFactory.h:108-110; ignored unless FM */
  uint32_t reserved0;
  const float* audio_taps; /* HOST pointer or NULL: no audio FIR (output = demodulated samples) */
  size_t audio_tap_count;
  size_t audio_decimation;
  int32_t cuda_device;
  uint32_t reserved1;
} b200sdr_chain_config;

typedef struct b200sdr_chain b200sdr_chain;

/* ---- life cycle ------------------------------------------------------------------------------- */
B200SDR_EXPORT b200sdr_status b200sdr_chain_create(const b200sdr_chain_config* config, b200sdr_chain** chainOut);
B200SDR_EXPORT void b200sdr_chain_destroy(b200sdr_chain* chain);
B200SDR_EXPORT const char* b200sdr_last_error(void); /* thread-local, never NULL */

/* ---- sample-count rules (pure host arithmetic; usable without a GPU) --------------------------- */
/* Fir::getNumOutputElements (Fir.cpp:141-187) in its well-defined form: nIn+1 >= T ? (nIn+1-T)/D : 0 */
B200SDR_EXPORT size_t b200sdr_fir_num_outputs(size_t numInputs, size_t tapCount, size_t decimation);
/* Outputs of each stage for numInputs input samples: RF FIR, demodulator (FM keeps one sample of
 * history, QuadFmDemod.cpp:76-84), audio FIR.  Any out pointer may be NULL. */
B200SDR_EXPORT void b200sdr_chain_counts(
    const b200sdr_chain* chain, size_t numInputs, size_t* numRf, size_t* numDemod, size_t* numAudio);
/* Input samples that one block producing numAudio final outputs reads, and the input stride between
 * consecutive final outputs (D1*D2).  span(numAudio) = (numAudio-1)*stride + window. */
B200SDR_EXPORT size_t b200sdr_chain_input_stride(const b200sdr_chain* chain);
B200SDR_EXPORT size_t b200sdr_chain_input_window(const b200sdr_chain* chain);
/* 64-bit fixed-point mixer phase increment: round(frac(frequency/sampleRate) * 2^64). */
B200SDR_EXPORT uint64_t b200sdr_phase_step(double frequency, double sampleRate);

/* Split `numAudio` final outputs into `parts` contiguous segments (time-segment sharding over GPUs or
 * over staged host chunks).  For part `index` returns the first output, the output count, the first
 * input sample it reads and the number of input samples it reads (overlap = look-ahead halo). */
B200SDR_EXPORT b200sdr_status b200sdr_chain_segment(
    const b200sdr_chain* chain, size_t numAudio, size_t parts, size_t index, size_t* firstOutput, size_t* outputCount,
    size_t* firstInput, size_t* inputCount);

/* The same with the outputs split in proportion to `weights` (HOST, `parts` positive numbers, e.g. each GPU's measured rate):
 * on a box whose GPUs do not run at the same speed (power capping spreads them by a few per cent) the faster devices get the
 * longer segments and all finish together.  weights == NULL: equal shares, as b200sdr_chain_segment. */
B200SDR_EXPORT b200sdr_status b200sdr_chain_segment_weighted(
    const b200sdr_chain* chain, size_t numAudio, size_t parts, const double* weights, size_t index, size_t* firstOutput,
    size_t* outputCount, size_t* firstInput, size_t* inputCount);

/* ---- device-resident processing (pointers are DEVICE pointers; async on `stream`) ---------------- */
/* Stage K1: convert + mix + decimating FIR + demod in one kernel.  `numOutputs` demodulated floats
 * (AM/FM) or complex RF-FIR samples (NONE) are written to `output`.  `firstSampleIndex` is the absolute
 * index of input[0] (mixer phase; irrelevant to AM/FM results).  `numInputs` bounds the reads. */
B200SDR_EXPORT b200sdr_status b200sdr_chain_rf_stage(
    b200sdr_chain* chain, const void* input, size_t numInputs, uint64_t firstSampleIndex, void* output,
    size_t numOutputs, cudaStream_t stream);
/* Stage K2: audio FIR (float taps, float data, decimating). */
B200SDR_EXPORT b200sdr_status b200sdr_chain_audio_stage(
    b200sdr_chain* chain, const float* demod, float* audio, size_t numAudio, cudaStream_t stream);
/* The whole chain over one block, producing EXACTLY numAudio final outputs (time-segment sharding, host staging).
 * Needs numInputs >= (numAudio-1)*stride + window.  When the shape allows (AM/FM with an audio FIR, D1 a multiple
 * of the 16-byte vector, <= 8 taps per decimation phase, 16-byte aligned input) this is ONE persistent kernel and
 * `demodScratch` is not touched (may be NULL); otherwise K1 + K2 through demodScratch. */
B200SDR_EXPORT b200sdr_status b200sdr_chain_run(
    b200sdr_chain* chain, const void* input, size_t numInputs, uint64_t firstSampleIndex, float* demodScratch,
    float* audio, size_t numAudio, cudaStream_t stream);
/* Same, with the number of outputs given by the reference's count rules (b200sdr_chain_counts).  `demodScratch` must hold numDemod floats (see b200sdr_chain_counts).
 * *numAudioOut receives the number of final outputs written to `audio`. */
B200SDR_EXPORT b200sdr_status b200sdr_chain_process_device(
    b200sdr_chain* chain, const void* input, size_t numInputs, uint64_t firstSampleIndex, float* demodScratch,
    float* audio, size_t audioCapacity, size_t* numAudioOut, cudaStream_t stream);

/* ---- host-buffer processing (the end-to-end path) ------------------------------------------------ */
/* Processes a block that lives in HOST memory (pinned for full PCIe rate) and returns the final
 * outputs in HOST memory.  Internally: the block is cut into overlapped time segments that are staged
 * through double-buffered device memory (H2D copy of segment i+1 overlaps K1/K2 of segment i and the D2H
 * of segment i-1).  Synchronous: returns when `audio` is complete. */
B200SDR_EXPORT b200sdr_status b200sdr_chain_process_host(
    b200sdr_chain* chain, const void* hostInput, size_t numInputs, uint64_t firstSampleIndex, float* hostAudio,
    size_t audioCapacity, size_t* numAudioOut);
/* Staging segment size (input samples) used by b200sdr_chain_process_host; default 2^25. */
B200SDR_EXPORT b200sdr_status b200sdr_chain_set_host_segment(b200sdr_chain* chain, size_t inputSamples);

/* ---- wideband channelizer (BASELINE config C5) --------------------------------------------------- */
/* Many channels out of ONE wideband int8 IQ stream: per channel mix by its own frequency -> shared RF low-pass
 * (decimating) -> AM or FM demodulation -> shared audio FIR (decimating).  Equivalent to num_channels instances of the
 * chain above (i.e. of the reference graph src/filters/factories/RfToPcmAudioFactory.cpp:214-304 behind one
 * Int8ToFloat) fed with the same input, but the input is read once per group of four channels and the mix + FIR runs
 * as an int8 GEMM on the tensor cores.  Multi-GPU: every rank builds a channelizer over ITS channels (sharding by
 * channel, no collective on the filter path). */
typedef struct b200sdr_channelizer_config {
  uint32_t struct_size;
  uint32_t num_channels;
  double sample_rate;             /* Hz, of the input */
  const double* frequencies;      /* HOST, per channel: the cosine-source frequency (tuned - channel) */
  const uint32_t* modulations;    /* HOST, per channel: B200SDR_MOD_AM / B200SDR_MOD_FM */
  const float* fm_gains;          /* HOST, per channel; ignored for AM channels */
  const float* rf_taps;           /* HOST, correlation order; shared by all channels */
  size_t rf_tap_count;
  size_t rf_decimation;           /* must be a multiple of 8 (16-byte input pieces); ceil(taps / decimation) <= 8 */
  const float* audio_taps;        /* HOST; shared by all channels */
  size_t audio_tap_count;
  size_t audio_decimation;
  int32_t cuda_device;
  uint32_t reserved;
} b200sdr_channelizer_config;

typedef struct b200sdr_channelizer b200sdr_channelizer;

B200SDR_EXPORT b200sdr_status b200sdr_channelizer_create(const b200sdr_channelizer_config* config, b200sdr_channelizer** out);
B200SDR_EXPORT void b200sdr_channelizer_destroy(b200sdr_channelizer* channelizer);
/* Output counts COMMON to all channels for numInputs input samples (the reference's count rules, as b200sdr_chain_counts;
 * with FM channels in the set this is the FM count -- see b200sdr_channelizer_channel_counts for a channel's own). */
B200SDR_EXPORT void b200sdr_channelizer_counts(const b200sdr_channelizer* channelizer, size_t numInputs, size_t* numDemod, size_t* numAudio);
/* input: DEVICE int8 IQ (16-byte aligned); demodScratch: DEVICE, num_channels * demodStride floats with
 * demodStride >= numDemod needed for numAudio outputs ((numAudio-1)*D2 + T2); audio: DEVICE, [channel][audioStride].
 * Launches: one RF kernel over all channel groups + one batched audio FIR. */
B200SDR_EXPORT b200sdr_status b200sdr_channelizer_run(
    b200sdr_channelizer* channelizer, const void* input, size_t numInputs, float* demodScratch, size_t demodStride, float* audio,
    size_t audioStride, size_t numAudio, cudaStream_t stream);
/* Output counts of ONE channel by the reference's per-node rules (Fir.cpp:141-187; QuadFmDemod.cpp:76-84 holds one sample
 * back, QuadAmDemod.cpp:80-107 none): an AM channel next to FM channels owns its single-chain count, which may be one audio
 * sample more than b200sdr_channelizer_counts (the count common to all channels) reports. */
B200SDR_EXPORT b200sdr_status b200sdr_channelizer_channel_counts(
    const b200sdr_channelizer* channelizer, uint32_t channel, size_t numInputs, size_t* numDemod, size_t* numAudio);
/* One block by the reference's count rules, channel by channel: as b200sdr_channelizer_run with the common count, plus --
 * when the set mixes AM and FM and the AM channels own one more output -- an AM pass over the input window of that last
 * output.  numAudioPerChannel (HOST, num_channels entries, may be NULL) receives each channel's count = what
 * b200sdr_channelizer_channel_counts reports.  audioStride must hold the largest count, demodStride >= max(numDemod needed,
 * audio_tap_count). */
B200SDR_EXPORT b200sdr_status b200sdr_channelizer_process(
    b200sdr_channelizer* channelizer, const void* input, size_t numInputs, float* demodScratch, size_t demodStride, float* audio,
    size_t audioStride, size_t* numAudioPerChannel, cudaStream_t stream);
B200SDR_EXPORT const char* b200sdr_channelizer_variant(const b200sdr_channelizer* channelizer);
/* The coarsest raster the channel set lies on: the smallest N in {4, 8, ..., 256} with frequencies[c] = frequencies[0] +
   b_c * sampleRate / N (decided on the 64-bit phase steps, i.e. mod sampleRate), or 0 if there is none.  bins (may be NULL)
   receives b_c mod N.  A non-zero answer is what sends b200sdr_channelizer_create down the filter-bank route (which may
   still pick a finer raster if the tables of the coarsest do not fit in shared memory).  Needs no GPU. */
B200SDR_EXPORT uint32_t b200sdr_channelizer_raster(const double* frequencies, uint32_t numChannels, double sampleRate, int32_t* bins);

/* Time-segment sharding of the channelizer (the filter-bank route yields every channel from one pass, so GPUs split the
 * audio outputs, not the channels): part `index` of `parts` of numAudio outputs per channel -- first output, count, first
 * input sample and number of input samples it reads (look-ahead halo included).  As b200sdr_chain_segment. */
B200SDR_EXPORT b200sdr_status b200sdr_channelizer_segment(
    const b200sdr_channelizer* channelizer, size_t numAudio, size_t parts, size_t index, size_t* firstOutput, size_t* outputCount,
    size_t* firstInput, size_t* inputCount);

/* ---- multi-GPU: the gather of decimated audio to rank 0 (the ONLY exchange of the sharded path) ------------------- */
/* One rank per GPU (one process per GPU, or one host thread per GPU).  Every rank runs its own time segment
 * (b200sdr_chain_segment / b200sdr_channelizer_segment) or its own channels -- no collective on the filter path
 * (reference: none; src/commandqueue/CommandQueueFactory.cpp:59-62 only lets a queue name a device).  The audio is
 * written into a slab of this object and gathered to rank 0 with one grouped NCCL send/recv per slab on a side stream,
 * overlapping the kernels of the next steps.  NCCL is loaded at run time (libnccl.so.2); world == 1 needs none. */
typedef struct b200sdr_gather_config {
  uint32_t struct_size;
  int32_t rank, world;
  uint32_t slabs;                 /* >= 2; 3 lets the main stream run ahead while two gathers drain */
  int32_t cuda_device;
  const size_t* floats_per_rank;  /* HOST, `world` entries: capacity of each rank's part of one slab, in floats */
  const void* nccl_unique_id;     /* NCCL mode: 128 bytes made by b200sdr_nccl_unique_id() on rank 0, handed to every rank by the caller */
  uint32_t mode;                  /* B200SDR_GATHER_NCCL, B200SDR_GATHER_PEER or B200SDR_GATHER_PEER_COPY */
  uint32_t reserved;
} b200sdr_gather_config;
/* NCCL: one grouped ncclSend/ncclRecv per slab.  PEER: rank 0's gathered slabs are mapped into every rank of the box (CUDA IPC
 * over NVLink / NVSwitch peer memory) and b200sdr_gather_slab() returns THAT memory, so the filter kernels store their audio
 * straight into rank 0's buffer.  PEER_COPY: the same mapping, but b200sdr_gather_slab() returns a local slab and
 * b200sdr_gather_submit() moves it into rank 0's memory with a copy engine (cudaMemcpyAsync over NVLink) on the side stream.
 * Neither peer mode costs a kernel or an SM; completion and reuse are sequenced by stream-ordered 32-bit flags.  The peer modes
 * need one exchange of handles after create: every rank exports a blob of b200sdr_gather_exchange_size() bytes, the caller
 * all-gathers the blobs (rank order) and every rank imports them. */
enum { B200SDR_GATHER_NCCL = 0, B200SDR_GATHER_PEER = 1, B200SDR_GATHER_PEER_COPY = 2 };
typedef struct b200sdr_gather b200sdr_gather;
B200SDR_EXPORT b200sdr_status b200sdr_nccl_unique_id(void* id128);
B200SDR_EXPORT b200sdr_status b200sdr_gather_create(const b200sdr_gather_config* config, b200sdr_gather** out);
B200SDR_EXPORT void b200sdr_gather_destroy(b200sdr_gather* gather);
B200SDR_EXPORT size_t b200sdr_gather_exchange_size(const b200sdr_gather* gather);
B200SDR_EXPORT b200sdr_status b200sdr_gather_export(b200sdr_gather* gather, void* blob);
B200SDR_EXPORT b200sdr_status b200sdr_gather_import(b200sdr_gather* gather, const void* blobsOfAllRanks);
/* DEVICE pointer of this rank's part of slab `slab` (floats_per_rank[rank] floats): the kernels write their audio here. */
B200SDR_EXPORT float* b200sdr_gather_slab(b200sdr_gather* gather, uint32_t slab);
/* Before writing into a slab again: `stream` waits until the slab's previous gather has read it. */
B200SDR_EXPORT b200sdr_status b200sdr_gather_acquire(b200sdr_gather* gather, uint32_t slab, cudaStream_t stream);
/* After the work that fills the slab was enqueued on `stream`: gather it on the side stream.  floatsPerRank (HOST, `world`
 * entries, identical on every rank; NULL = the capacities) gives what each rank really sends. */
B200SDR_EXPORT b200sdr_status b200sdr_gather_submit(b200sdr_gather* gather, uint32_t slab, const size_t* floatsPerRank, cudaStream_t stream);
/* `stream` waits for every gather submitted so far. */
B200SDR_EXPORT b200sdr_status b200sdr_gather_finish(b200sdr_gather* gather, cudaStream_t stream);
/* Rank 0: DEVICE pointer of rank `rank`'s part of gathered slab `slab` (valid once the gather has run); NULL elsewhere. */
B200SDR_EXPORT const float* b200sdr_gather_result(const b200sdr_gather* gather, uint32_t slab, int32_t rank);
B200SDR_EXPORT void b200sdr_gather_stats(const b200sdr_gather* gather, uint64_t* gathers, uint64_t* floatsMoved, int32_t* ncclVersion);

/* ---- introspection ------------------------------------------------------------------------------ */
/* Kernels launched by this library since load (all streams); used by bench.py's gpu_launches. */
B200SDR_EXPORT uint64_t b200sdr_launch_count(void);
/* Name of the kernel variant K1 resolves to for this chain ("rows<MP=4,RPT=4,...>" or "direct"). */
B200SDR_EXPORT const char* b200sdr_chain_variant(const b200sdr_chain* chain);
B200SDR_EXPORT const char* b200sdr_version(void);
/* Host-side tables of the Toeplitz chain kernel (csrc/toeplitz_kernels.cuh) for an int8 chain with these RF taps, decimation
   (a multiple of 8) and mixer: the three-int8-digit fragments of B in the order the kernel reads them
   ([pair of k-steps q][lane 32][(ksub*3 + digit)*2 + half], one 32-bit word = four k-rows), the digit weights and the
   number of k-steps.  Needs no GPU; tests/test_toeplitz_tables.py rebuilds B from it and runs the contraction in numpy.
   fragWords receives the word count; fragments may be NULL to query it. */
B200SDR_EXPORT b200sdr_status b200sdr_toeplitz_tables(
    const float* rfTaps, size_t rfTapCount, size_t rfDecimation, uint32_t mix, double frequency, double sampleRate,
    uint32_t* fragments, size_t fragCapacityWords, size_t* fragWords, float digitScale[3], uint32_t* kSteps, uint32_t* magic);

#endif /* B200SDR_B200SDR_H */
