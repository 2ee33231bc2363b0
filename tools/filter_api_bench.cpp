// filter_api_bench -- end-to-end throughput THROUGH THE REFERENCE-FACING PLUGIN API (getFactoriesSingleton(), Filter::
// requestBuffer / commitBuffer / readOutput), host buffers in and out:
//
//   producer (host threads) -> CudaMemcpyFilter H2D (pinned, two blocks) -> ONE fused chain Filter -> CudaMemcpyFilter D2H
//                           -> two alternating pinned result buffers behind an IEventPipeline (the reference's Waiter)
//
// This is the `e2e` leg of bench.py: the same Sink/Source calls an application written against the reference makes
// (src/applications/nbfm_test.cpp:256-354), with the five-node chain replaced by the fused node.  The producer stands in
// for a capture source (HackrfSource / FileSource reading a recording from the page cache): a pool of host threads copies the
// next block of a pre-generated synthetic int8 IQ capture STRAIGHT into the copy node's pinned block (requestBuffer -> write
// -> commitBuffer), so the only host copy on the path is the producer's own write.
//
// usage: filter_api_bench --taps1 F32 --taps2 F32 [--fs HZ --freq HZ --mod am|fm --dev HZ --d1 N --d2 N] [--device I]
//          [--samples-per-pass N --passes P --step BYTES --warmup-steps W --pipeline 0|1 --threads T --producer copy|resident]
// --producer resident models a capture device that DMAs into the pinned block itself: a block is written the first time the
// copy node hands it out and committed as it is afterwards (no host copy in the timed region; the H2D copy stays).
// Prints one JSON line.
#include <cuda_runtime.h>
#include <gpusdrpipeline/EventPipeline.h>
#include <gpusdrpipeline/Factories.h>
#include <gpusdrpipeline/FusedChain.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace std;

static vector<float> readFloats(const string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    fprintf(stderr, "cannot open %s\n", path.c_str());
    exit(2);
  }
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  vector<float> v(static_cast<size_t>(n) / sizeof(float));
  if (n > 0 && fread(v.data(), sizeof(float), v.size(), f) != v.size()) exit(2);
  fclose(f);
  return v;
}

// xorshift64* bytes = full-range int8 IQ noise
static void fillNoise(uint8_t* dst, size_t lo, size_t hi, uint64_t seed) {
  uint64_t x = seed ^ (lo * 0x9E3779B97F4A7C15ull) ^ 0x2545F4914F6CDD1Dull;
  size_t i = lo;
  for (; i + 8 <= hi; i += 8) {
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    const uint64_t v = x * 0x2545F4914F6CDD1Dull;
    memcpy(dst + i, &v, 8);
  }
  for (; i < hi; i++) dst[i] = static_cast<uint8_t>(x >> (8 * (i & 7)));
}

// A fixed pool of producer threads (created once: a block is ~1 ms of work, thread start-up would show).  run(n, fn) calls
// fn(lo, hi) on the threads' shares of [0, n) and returns when all are done.
class Pool {
 public:
  explicit Pool(unsigned threads) : mCount(threads ? threads : 1) {
    for (unsigned t = 1; t < mCount; t++) mThreads.emplace_back([this, t] { worker(t); });
  }
  ~Pool() {
    {
      lock_guard<mutex> lock(mMutex);
      mStop = true;
      mGeneration++;
    }
    mWake.notify_all();
    for (auto& th : mThreads) th.join();
  }
  void run(size_t n, const function<void(size_t, size_t)>& fn) {
    if (mCount == 1 || n <= (size_t(1) << 20)) {
      fn(0, n);
      return;
    }
    {
      lock_guard<mutex> lock(mMutex);
      mFn = &fn;
      mN = n;
      mPending = mCount - 1;
      mGeneration++;
    }
    mWake.notify_all();
    share(0);
    unique_lock<mutex> lock(mMutex);
    mDone.wait(lock, [this] { return mPending == 0; });
  }

 private:
  void share(unsigned t) {
    const size_t per = ((mN + mCount - 1) / mCount + 63) & ~size_t(63);
    const size_t lo = t * per, hi = lo + per < mN ? lo + per : mN;
    if (lo < hi) (*mFn)(lo, hi);
  }
  void worker(unsigned t) {
    uint64_t seen = 0;
    for (;;) {
      {
        unique_lock<mutex> lock(mMutex);
        mWake.wait(lock, [&] { return mGeneration != seen; });
        seen = mGeneration;
        if (mStop) return;
      }
      share(t);
      {
        lock_guard<mutex> lock(mMutex);
        if (--mPending == 0) mDone.notify_one();
      }
    }
  }
  unsigned mCount;
  vector<thread> mThreads;
  mutex mMutex;
  condition_variable mWake, mDone;
  const function<void(size_t, size_t)>* mFn = nullptr;
  size_t mN = 0;
  unsigned mPending = 0;
  uint64_t mGeneration = 0;
  bool mStop = false;
};

int main(int argc, char** argv) {
  double fs = 19.2e6, freq = -1.234e6, dev = 75e3;
  string mod = "am", taps1Path, taps2Path;
  size_t d1 = 40, d2 = 10, samplesPerPass = size_t(1) << 28, passes = 4, step = size_t(64) << 20, warmupSteps = 4;
  int device = 0;
  bool pipelined = true;
  unsigned threads = 0;
  bool resident = false;
  for (int i = 1; i + 1 < argc; i += 2) {
    const string k = argv[i], v = argv[i + 1];
    if (k == "--fs") fs = atof(v.c_str());
    else if (k == "--freq") freq = atof(v.c_str());
    else if (k == "--dev") dev = atof(v.c_str());
    else if (k == "--mod") mod = v;
    else if (k == "--d1") d1 = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--d2") d2 = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--taps1") taps1Path = v;
    else if (k == "--taps2") taps2Path = v;
    else if (k == "--device") device = atoi(v.c_str());
    else if (k == "--samples-per-pass") samplesPerPass = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--passes") passes = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--step") step = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--warmup-steps") warmupSteps = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--pipeline") pipelined = v != "0";
    else if (k == "--threads") threads = static_cast<unsigned>(strtoul(v.c_str(), nullptr, 10));
    else if (k == "--producer") resident = v == "resident";
    else {
      fprintf(stderr, "unknown argument %s\n", k.c_str());
      return 2;
    }
  }
  if (taps1Path.empty() || taps2Path.empty()) {
    fprintf(stderr, "--taps1 and --taps2 are required\n");
    return 2;
  }
  if (threads == 0) threads = thread::hardware_concurrency() ? thread::hardware_concurrency() : 1;
  if (threads > 32) threads = 32;
  const vector<float> taps1 = readFloats(taps1Path), taps2 = readFloats(taps2Path);
  const bool fm = mod == "fm";
  gslogSetVerbosity(GSLOG_WARN);

  ConstRef<IFactories> factories = unwrap(getFactoriesSingleton());
  ConstRef<ICudaCommandQueue> queue = unwrap(factories->getCudaCommandQueueFactory()->create(device));
  ConstRef<Filter> h2d = unwrap(factories->getCudaMemcpyFilterFactory()->createCudaMemcpy(cudaMemcpyHostToDevice, queue));
  ConstRef<Filter> d2h = unwrap(factories->getCudaMemcpyFilterFactory()->createCudaMemcpy(cudaMemcpyDeviceToHost, queue));
  GsFusedChainParams p {};
  p.structSize = sizeof(p);
  p.inputType = SampleType_Int8Complex;
  p.modulation = fm ? Modulation_Fm : Modulation_Am;
  p.mix = 1;
  p.sampleRate = fs;
  p.frequency = freq;
  p.rfTaps = taps1.data();
  p.rfTapCount = taps1.size();
  p.rfDecimation = d1;
  p.fmGain = static_cast<float>(fs / static_cast<double>(d1)) / (2.0f * static_cast<float>(M_PI) * static_cast<float>(dev) * 5);
  p.audioTaps = taps2.data();
  p.audioTapCount = taps2.size();
  p.audioDecimation = d2;
  ConstRef<Filter> chain = unwrap(gsCreateFusedChain(&p, queue));
  ConstRef<IAllocator> pinnedAlloc = unwrap(factories->getCudaAllocatorFactory()->createCudaAllocator(queue, 32, true));
  ConstRef<IBufferFactory> pinnedFactory = unwrap(factories->createBufferFactory(pinnedAlloc));
  const size_t resultBytes = step / 2 / (d1 * d2) * sizeof(float) * 2 + (1 << 16);
  ConstRef<IBuffer> host[2] = {unwrap(pinnedFactory->createBuffer(resultBytes)), unwrap(pinnedFactory->createBuffer(resultBytes))};
  Ref<IEventPipeline> pipeline;
  if (pipelined) pipeline = unwrap(gsCreateEventPipeline(queue));

  const size_t passBytes = samplesPerPass * 2;
  // the "capture": two blocks of noise; every step copies one block's worth from a different offset
  Pool pool(threads);
  vector<uint8_t*> filledBlocks;  // --producer resident: pinned blocks that already hold a full block of samples
  vector<uint8_t> capture(2 * step + 64);
  pool.run(capture.size(), [&](size_t lo, size_t hi) { fillNoise(capture.data(), lo, hi, 0x9E3779B97F4A7C15ull); });
  size_t total = 0, timedFrom = 0, outputs = 0, stepNo = 0, h2dBytes = 0, d2hBytes = 0;
  double checksum = 0.0;
  int pendingSlot = -1;  // result buffer whose copy was enqueued one step ago and has not been read yet
  auto harvest = [&](int slot) {  // the device->host copy into host[slot] has completed: the host reads the audio
    const size_t n = host[slot]->range()->used() / sizeof(float);
    const float* a = host[slot]->readPtr<float>();
    for (size_t i = 0; i < n; i += 64) checksum += a[i];
    outputs += n;
    d2hBytes += n * sizeof(float);
  };
  auto sync = [&] {
    cudaSetDevice(queue->cudaDevice());
    cudaStreamSynchronize(queue->cudaStream());
  };
  IBuffer* o[1];
  // host time per phase of the loop (timed region only), printed with the result: where the producer thread waits
  double tRequest = 0, tProduce = 0, tEnqueue = 0, tWait = 0;
  auto now = [] { return chrono::steady_clock::now(); };
  auto since = [&](chrono::steady_clock::time_point t0) { return chrono::duration<double>(now() - t0).count(); };
  auto start = chrono::steady_clock::now();
  for (size_t pass = 0; pass < passes; pass++) {
    for (size_t pos = 0; pos < passBytes; stepNo++) {
      if (stepNo == warmupSteps && stepNo > 0) {  // pinned blocks and port buffers are allocated in the first steps
        if (pendingSlot >= 0) {
          THROW_IF_ERR(pipeline->waitLast());
          harvest(pendingSlot);
          pendingSlot = -1;
        }
        sync();
        timedFrom = total;
        h2dBytes = d2hBytes = 0;
        tRequest = tProduce = tEnqueue = tWait = 0;
        start = chrono::steady_clock::now();
      }
      size_t bytes = passBytes - pos < step ? passBytes - pos : step;
      bytes &= ~size_t(1);
      auto t0 = now();
      Ref<IBuffer> staged = unwrap(h2d->requestBuffer(0, bytes));
      tRequest += since(t0);
      t0 = now();
      uint8_t* const block = staged->writePtr();
      bool written = false;
      for (uint8_t* seen : filledBlocks) written = written || seen == block;
      if (!resident || !written) {
        if (!written && filledBlocks.size() < 16) filledBlocks.push_back(block);
        uint8_t* dst = block;
        const uint8_t* src = capture.data() + (stepNo * 4099 * 2) % step;  // whole samples, a different offset every step
        pool.run(bytes, [=](size_t lo, size_t hi) { memcpy(dst + lo, src + lo, hi - lo); });
      }
      tProduce += since(t0);
      t0 = now();
      THROW_IF_ERR(h2d->commitBuffer(0, bytes));
      pos += bytes;
      total += bytes / 2;
      h2dBytes += bytes;
      size_t want = h2d->getAlignedOutputDataSize(0);
      Ref<IBuffer> chainIn = unwrap(chain->requestBuffer(0, want < 256 ? 256 : want));
      o[0] = chainIn.get();
      THROW_IF_ERR(h2d->readOutput(o, 1));
      THROW_IF_ERR(chain->commitBuffer(0, chainIn->range()->used()));
      want = chain->getAlignedOutputDataSize(0);
      Ref<IBuffer> d2hIn = unwrap(d2h->requestBuffer(0, want < 256 ? 256 : want));
      o[0] = d2hIn.get();
      THROW_IF_ERR(chain->readOutput(o, 1));
      THROW_IF_ERR(d2h->commitBuffer(0, d2hIn->range()->used()));
      const int slot = static_cast<int>(stepNo & 1);
      host[slot]->range()->clearRange();
      o[0] = host[slot].get();
      THROW_IF_ERR(d2h->readOutput(o, 1));
      if (d2h->getOutputDataSize(0) != 0) {
        fprintf(stderr, "result buffer too small\n");
        return 3;
      }
      tEnqueue += since(t0);
      t0 = now();
      if (pipelined) {
        // wait for the PREVIOUS step only (src/filters/Waiter.cpp:34-50): this step's copies and kernel keep running
        THROW_IF_ERR(pipeline->recordNextAndWaitPrevious());
        if (pendingSlot >= 0) harvest(pendingSlot);
        pendingSlot = slot;
      } else {
        sync();  // a stream synchronisation per step, as nbfm_test.cpp:346-347 does
        harvest(slot);
      }
      tWait += since(t0);
    }
  }
  if (pendingSlot >= 0) {
    THROW_IF_ERR(pipeline->waitLast());
    harvest(pendingSlot);
  }
  sync();
  const double secs = chrono::duration<double>(chrono::steady_clock::now() - start).count();
  const size_t timedSteps = stepNo > warmupSteps ? stepNo - warmupSteps : stepNo;
  printf("{\"samples\": %zu, \"timed_samples\": %zu, \"timed_steps\": %zu, \"outputs\": %zu, \"seconds\": %.6f, \"msps\": %.3f, \"step_bytes\": %zu, "
         "\"h2d_bytes\": %zu, \"d2h_bytes\": %zu, \"pipeline\": %s, \"threads\": %u, \"device\": %d, \"producer\": \"%s\", "
         "\"host_ms_per_step\": {\"request\": %.4f, \"produce\": %.4f, \"enqueue\": %.4f, \"wait\": %.4f}, \"checksum\": %.6g}\n",
         total, total - timedFrom, timedSteps, outputs, secs, static_cast<double>(total - timedFrom) / secs / 1e6, step, h2dBytes, d2hBytes,
         pipelined ? "true" : "false", threads, device, resident ? "resident" : "copy", 1e3 * tRequest / timedSteps, 1e3 * tProduce / timedSteps, 1e3 * tEnqueue / timedSteps,
         1e3 * tWait / timedSteps, checksum);
  return 0;
}
