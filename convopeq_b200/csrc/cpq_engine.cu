// cpq_engine.cu -- engine + C ABI (include/cpq.h) over the sm_100a kernels in cpq_kernels.cuh.
//
// Offline form of the reference's per-callback chain
//   convolverRt().process -> eqRt().process(block, params, cache) -> makeup gain -> processOutputDouble
// (AudioEngine.Processing.DSPCoreDouble.cpp:386-414,465-469,577-663) for a batch of independent streams.
// There is no CPU path: every compute entry point needs a CUDA device.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/cpq.h"
#include "cpq_plan.hpp"
#include "cpq_fft.cuh"
#include "cpq_fft16.cuh"
#include "cpq_fft_large.cuh"
#include "cpq_mac.cuh"
#include "cpq_eq.cuh"

namespace cpq
{

#define CPQ_CUDA(expr)                                                                              \
    do                                                                                              \
    {                                                                                               \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
        {                                                                                           \
            setError(std::string(#expr) + ": " + cudaGetErrorString(e_));                           \
            return e_ == cudaErrorMemoryAllocation ? CPQ_ERR_OOM : CPQ_ERR_CUDA;                    \
        }                                                                                           \
    } while (0)

// Guard mode (CPQ_GUARD=1 in the environment, a debugging aid: compute-sanitizer is not available on every pool): every
// device buffer of the engine is allocated with a 256-byte canary before and after it; cpq_debug_check_guards() reports how
// many canaries a kernel has overwritten.  tests/test_guards.py runs every kernel family under it.
static const bool g_guard = [] { const char* e = getenv("CPQ_GUARD"); return e && atoi(e) != 0; }();
constexpr size_t kGuardBytes = 256;
struct GuardRegistry
{
    std::vector<std::pair<unsigned char*, size_t>> live;   // raw allocation, payload bytes
    static GuardRegistry& get() { static GuardRegistry r; return r; }
};

template <typename T>
struct DevBuf
{
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release()
    {
        if (p)
        {
            if (g_guard)
            {
                unsigned char* raw = reinterpret_cast<unsigned char*>(p) - kGuardBytes;
                auto& v = GuardRegistry::get().live;
                for (size_t i = 0; i < v.size(); ++i)
                    if (v[i].first == raw) { v.erase(v.begin() + (long) i); break; }
                cudaFree(raw);
            }
            else
                cudaFree(p);
        }
        p = nullptr;
        n = 0;
    }
    cudaError_t ensure(size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        release();
        const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        if (!g_guard)
        {
            cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), bytes);
            if (e == cudaSuccess) n = count;
            return e;
        }
        const size_t padded = (bytes + 15) / 16 * 16;
        unsigned char* raw = nullptr;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&raw), padded + 2 * kGuardBytes);
        if (e != cudaSuccess) return e;
        cudaMemset(raw, 0xA5, kGuardBytes);
        cudaMemset(raw + kGuardBytes + padded, 0xA5, kGuardBytes);
        GuardRegistry::get().live.push_back({ raw, padded });
        p = reinterpret_cast<T*>(raw + kGuardBytes);
        n = count;
        return cudaSuccess;
    }
};

// Page-locked host memory (staging slots for pageable caller buffers)
struct PinnedBuf
{
    double* p = nullptr;
    size_t n = 0;
    ~PinnedBuf() { release(); }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t ensure(size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&p), std::max<size_t>(count, 1) * sizeof(double), cudaHostAllocDefault);
        if (e == cudaSuccess) n = count;
        return e;
    }
};

// Pageable caller buffers (what a host application hands over: std::vector, juce::AudioBuffer, numpy).  cudaMemcpyAsync on
// such memory is staged by the driver through its own bounce buffer on the calling thread -- about 7 GB/s, H2D and D2H one after
// the other.  Instead the rows of sequence chunk c are copied by a few host threads into slot c % kSlots of a pinned ring, from
// where the chunk goes down as the pinned path's copies do; results come up into a second ring and are copied out by a second
// set of threads.  The calling thread only enqueues; it waits for "chunk staged" before it enqueues the chunk's H2D and for
// "slot drained" before it lets a D2H overwrite a slot.
struct HostStager
{
    static constexpr int kSlots = 3;
    struct Chunk { int s0 = 0, ns = 0; };
    int device = 0;
    int nThreads = 0;
    int64_t T = 0;
    size_t slotElems = 0;
    double* const* rows = nullptr;        // caller rows, indexed by absolute sequence
    double* in = nullptr;
    double* out = nullptr;
    std::vector<Chunk> chunks;
    std::vector<cudaEvent_t> evH2D, evD2H;   // recorded by the calling thread when it enqueues the copies
    std::vector<int> staged, drained;        // threads that have finished the chunk
    size_t h2dEnqueued = 0, d2hEnqueued = 0;
    bool abort = false;
    std::mutex m;
    std::condition_variable cv;
    std::vector<std::thread> threads;

    double* inSlot(size_t c) const { return in + (c % kSlots) * slotElems; }
    double* outSlot(size_t c) const { return out + (c % kSlots) * slotElems; }

    void start()
    {
        staged.assign(chunks.size(), 0);
        drained.assign(chunks.size(), 0);
        for (int j = 0; j < nThreads; ++j)
        {
            threads.emplace_back([this, j] { feed(j); });
            threads.emplace_back([this, j] { drain(j); });
        }
    }
    // thread j of the feeders: its share of the rows of every chunk, caller -> pinned slot
    void feed(int j)
    {
        cudaSetDevice(device);
        for (size_t c = 0; c < chunks.size(); ++c)
        {
            if (c >= (size_t) kSlots)
            {
                // the slot is free once the H2D of chunk c - kSlots has left it
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return abort || h2dEnqueued > c - kSlots; });
                if (abort) return;
                lk.unlock();
                cudaEventSynchronize(evH2D[c - kSlots]);
            }
            const Chunk& k = chunks[c];
            for (int r = j; r < k.ns; r += nThreads) std::memcpy(inSlot(c) + (size_t) r * T, rows[k.s0 + r], (size_t) T * sizeof(double));
            std::lock_guard<std::mutex> lk(m);
            if (abort) return;
            ++staged[c];
            cv.notify_all();
        }
    }
    void drain(int j)
    {
        cudaSetDevice(device);
        for (size_t c = 0; c < chunks.size(); ++c)
        {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return abort || d2hEnqueued > c; });
                if (abort) return;
            }
            cudaEventSynchronize(evD2H[c]);
            const Chunk& k = chunks[c];
            for (int r = j; r < k.ns; r += nThreads) std::memcpy(rows[k.s0 + r], outSlot(c) + (size_t) r * T, (size_t) T * sizeof(double));
            std::lock_guard<std::mutex> lk(m);
            if (abort) return;
            ++drained[c];
            cv.notify_all();
        }
    }
    // calling thread
    void waitStaged(size_t c)
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return staged[c] == nThreads; });
    }
    void waitDrained(size_t c)
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return drained[c] == nThreads; });
    }
    void noteH2D(size_t c)
    {
        std::lock_guard<std::mutex> lk(m);
        h2dEnqueued = c + 1;
        cv.notify_all();
    }
    void noteD2H(size_t c)
    {
        std::lock_guard<std::mutex> lk(m);
        d2hEnqueued = c + 1;
        cv.notify_all();
    }
    ~HostStager()
    {
        {
            std::lock_guard<std::mutex> lk(m);
            // a regular end has every chunk drained; anything else is an error exit: release the threads
            if (drained.empty() || drained.back() != nThreads) abort = true;
            cv.notify_all();
        }
        for (auto& t : threads) t.join();
    }
};

static bool hostRowIsPageable(const void* p)
{
    cudaPointerAttributes at {};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess)
    {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

// number of overwritten canaries among the live guarded buffers (0 = clean); -1 when guard mode is off
static int checkGuards()
{
    if (!g_guard) return -1;
    cudaDeviceSynchronize();
    int bad = 0;
    std::vector<unsigned char> h(kGuardBytes);
    for (auto& g : GuardRegistry::get().live)
        for (int side = 0; side < 2; ++side)
        {
            cudaMemcpy(h.data(), g.first + (side ? kGuardBytes + g.second : 0), kGuardBytes, cudaMemcpyDeviceToHost);
            for (unsigned char c : h)
                if (c != 0xA5) { ++bad; break; }
        }
    return bad;
}

struct EqSet
{
    cpq_svf_coeffs coeffs[CPQ_NUM_BANDS];
    uint8_t active[CPQ_NUM_BANDS];
    int32_t mode[CPQ_NUM_BANDS];
    double saturation = (double) 0.2f;
    double totalGain = 1.0;
    bool set = false;
    std::vector<GainEvent> events;
    GainRampState ramp;                      // streaming continuation: the total-gain ramp between calls
    int structure = 0;                       // EQParameters::filterStructure: 0 Serial, 1 Parallel
    int agc = 0;                             // EQParameters::agcEnabled
    uint8_t nodeActive[CPQ_NUM_BANDS];       // BandNode::active, used instead of `active` when the node path runs (Mid/Side)
    bool hasNodeActive = false;
};

struct LayerDev
{
    DevBuf<double2> H;      // [nH][Q][P] packed spectra
    DevBuf<double2> tw;     // [P+1]
    DevBuf<double2> ptw;    // per-pass radix-8 twiddle tables
    DevBuf<double> gain;    // [M] spectrum filter gain (when a FilterSpec is given)
    DevBuf<double> tilt;    // [M]
    bool hasGain = false, hasTilt = false;
    DevBuf<double2> X, Y;   // workspace spectra for a chunk of sequences
    DevBuf<double> tail;    // [chunk][K*P] (layers >= 1)
    DevBuf<int64_t> tailSrc;   // layers >= 1: delay-line position per callback; layer 0: ring position per callback (l0Src)
    DevBuf<int32_t> blockMap;
    DevBuf<int32_t> l0Count;   // layer 0 only: samples the ring delivers per callback (non-power-of-two host blocks)
    bool hasBlockMap = false;
};

struct Engine
{
    cpq_config cfg {};
    int nSeq = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    // convolver
    bool planSet = false;
    ConvPlan plan;
    std::vector<uint8_t> haveImpulse;   // per H row
    int nH = 0;
    LayerDev layer[CPQ_MAX_LAYERS];
    GatherPlan gplan;
    int64_t gplanCallbacks = -1;        // callbacks the device copies of the gather plan hold (one-shot calls)
    int64_t gplanHostCallbacks = -1;    // callbacks the host copy covers (streaming calls read their slice from it)
    int partBegin = 0, partEnd = -1;    // partition-range sharding
    bool outerPending = false;

    // io
    DevBuf<double> io;
    DevBuf<double> irScratch;

    // EQ
    std::vector<EqSet> eqSets;
    bool eqDirty = true;
    DevBuf<double> eqc, satDev, gainConst, gainTab, stateOut;
    DevBuf<unsigned> bandMask, scalarMask;
    DevBuf<int> setOfSeq;
    int64_t gainTabCallbacks = -1;
    bool haveGainTab = false;
    DevBuf<double> chainRec;
    DevBuf<unsigned> ticketFault;       // [0] ticket, [1] fault
    // EQ modes beyond Serial / Stereo|Left|Right
    bool anyPar = false, anyAgc = false, anyMs = false;
    unsigned msSplitMask = 0;                        // bands that are Mid/Side in at least one stream
    std::vector<int> msStreams[CPQ_NUM_BANDS];       // sorted streams whose band b is Mid/Side
    DevBuf<int> msStreamsDev[CPQ_NUM_BANDS], msSetDev[CPQ_NUM_BANDS];
    DevBuf<unsigned> msMaskDev[CPQ_NUM_BANDS];       // [2 * count]: bit b on the Mid row or on the Side row
    DevBuf<double> msScratch, sumsq, agcTab, agcState;
    // streaming continuation of Mid/Side bands: EQProcessor::filterState[2] / [3] (EQProcessor.h:637) -- per band region the
    // (Mid row, Side row) states of the streams that run the band there, in the order of msStreams[b]; region 20 = the rows of
    // Parallel-structure streams.  Zeroed by Reset and whenever the band settings are uploaded again.
    DevBuf<double> msState;
    bool msImported = false;
    // streaming continuation of the dry path (mix < 1 / bypass): the latency-compensation delay ring, ping-pong per call
    DevBuf<double> dryHist[2];
    int dryHistSel = 0, dryHistLen = 0;
    double* msRegion(int b) { return msState.p + (size_t) b * 2 * cfg.n_streams * CPQ_NUM_BANDS * 2; }
    size_t msStateCount() const { return (size_t) (CPQ_NUM_BANDS + 1) * 2 * cfg.n_streams * CPQ_NUM_BANDS * 2; }
    std::vector<int> parMsStreams;                   // sorted streams in the Parallel structure that have Mid/Side bands
    DevBuf<int> parMsStreamsDev, parMsSetDev;
    DevBuf<unsigned> parMsMaskDev;                   // [2 * count]: all Mid bands on the Mid row, all Side bands on the Side row, | 1 << 31
    DevBuf<double> parMsScratch;
    DevBuf<uint8_t> agcOnDev;
    cpq_status runEq(EqArgs e, int s0, int ns);

    // epilogue
    double makeup = 1.0;
    int ditherBits = 0;
    DevBuf<double> uniforms, ditherZ;
    int64_t uniformsPerCh = 0;

    // timing / pipelining
    cudaStream_t sIn = nullptr, sOut = nullptr;   // H2D and D2H streams of the host entry point
    static constexpr int kDitherStreams = 32;
    cudaStream_t sDither[kDitherStreams] {};      // the dither stage is serial in time (latency-bound): it runs beside the next chunks
    const double* uniformsBorrowed = nullptr;     // cpq_set_dither_uniforms_device
    bool ditherRng = false;                       // cpq_set_dither_seed: uniforms from the reference's fallback generator
    std::vector<unsigned long long> rngSeedState; // [nSeq] fallbackState right after construction
    DevBuf<unsigned long long> rngState;          // [nSeq] carried
    cudaEvent_t ev[8] {};
    std::vector<cudaEvent_t> evPool;
    cudaEvent_t poolEvent(size_t i)
    {
        while (evPool.size() <= i)
        {
            cudaEvent_t e = nullptr;
            cudaEventCreate(&e);
            evPool.push_back(e);
        }
        return evPool[i];
    }
    cpq_timings timings {};
    int64_t launches = 0;

    void setError(const std::string& s) { err = s; }

    ~Engine()
    {
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : evPool)
            if (e) cudaEventDestroy(e);
        for (auto& d : sDither)
            if (d) cudaStreamDestroy(d);
        if (sIn) cudaStreamDestroy(sIn);
        if (sOut) cudaStreamDestroy(sOut);
        if (stream) cudaStreamDestroy(stream);
    }

    int hRowOf(int stream_, int ch) const { return cfg.shared_ir ? ch : stream_ * cfg.n_channels + ch; }

    cpq_status init(const cpq_config* c);
    cpq_status setKernelAttributes();
    cpq_status setImpulse(int stream_, int ch, const double* ir, int len, double scale, const cpq_filter_spec* spec);
    cpq_status ensureTwiddles(int li);
    cpq_status uploadEq(int64_t nCallbacks);
    cpq_status uploadPost();
    struct OutCfg { int filterEnabled = 0, convIsLast = 0, hc = 1, lc = 0, lp = 1; double dcCutoff = 0.0; int finalClamp = 0; } outCfg;
    double convInputTrim = 1.0;         // state.convolverInputTrimGain (EQThenConvolver order only)
    double mix = 1.0;                   // (double) mixTarget of ConvolverProcessor::process (CPQ_CONV_OUTER only)
    int dryDelay = 0;                   // latency-compensation delay of the dry path, samples
    int winFirst = 0, winCount = -1;    // stream window of process calls (cpq_set_stream_window); -1 = all streams
    int nPeers = 0;                     // partition-range sharding: every rank's partial buffer, summed in the EQ launch's load
    const double* peers[CPQ_MAX_PEERS] {};
    int convBypassed = 0;               // runtimeSnapshot.bypassed: processBypassWithLatencyCompensation (Runtime.cpp:123-186)
    DevBuf<double> dryBuf;              // copy of the convolver input for the dry path (mix < 0.999) / the direct-form head
    double inputGain = 1.0;             // DSPCore::processInput's headroom gain (CPQ_STAGE_INPUT)
    double limiterMs = 0.0;             // SimplePeakLimiter release (ms); 0 = stage off.  The reference engine uses 100 ms
    DevBuf<unsigned> limFlag;           // [n_streams]
    DevBuf<double> limEnv;              // [n_streams] envelope after the last call
    int directHead = 0;                 // enableDirectHead of SetImpulse (experimental in the reference)
    DevBuf<double> directTaps;          // [nH][32] reversed, scaled head taps
    bool postDirty = true;
    unsigned postIdentity = 0;          // output-filter stages whose coefficients are the identity (skipped)
    DevBuf<double> postc, postState;
    cpq_status ensureGather(int64_t nCallbacks);
    cpq_status ensureGatherHost(int64_t nCallbacks);

    // ---- streaming continuation (cpq_set_streaming): the state the reference keeps between callbacks, carried between calls.
    // Per sequence: the last 2*Pmax input samples (prevInputBuf / inputAccBuf of every layer, MKLNonUniformConvolver.h:288-365),
    // per layer the last Q-1 input spectra (the FDL) and the tail-stream samples not yet read by Get (the delay line); the EQ,
    // output-stage, AGC, limiter and dither states live in the buffers the one-shot form already writes.
    bool streaming = false;
    bool contValid = false;              // false = the next call starts from Reset
    int64_t absCallback = 0;             // callbacks processed since Reset
    int histLen = 0;                     // input history samples per sequence
    int fdlRows[CPQ_MAX_LAYERS] = {};    // Q - 1
    int carryFrames[CPQ_MAX_LAYERS] = {};
    DevBuf<double> inHistBuf[2];         // input history, ping-pong per call (the forward transforms read it beside the call's input)
    int inHistSel = 0;
    double* inHistCur() { return inHistBuf[inHistSel].p; }
    // The FDL: per layer a persistent buffer of xCap rows of spectra per sequence.  Rows [xHead - fdlRows, xHead) are the spectra
    // of the Q - 1 frames before the next call's first one; a call appends its new frames at xHead (the forward transform writes
    // them there, the MAC reads [carried | new] in place) and xHead moves on.  When the buffer is full the carried rows are
    // copied back to its start -- once every (xCap - 2 fdlRows) / K calls instead of a copy in and a copy out per call.
    DevBuf<double2> xs[CPQ_MAX_LAYERS];
    int xCap[CPQ_MAX_LAYERS] = {}, xHead[CPQ_MAX_LAYERS] = {};
    cpq_status prepareFdl(int li, int64_t newFrames, bool cont);
    DevBuf<double> tailCarry[CPQ_MAX_LAYERS];
    cpq_status ensureStreamState();
    cpq_status resetState();
    size_t stateBytes() const;
    cpq_status exportState(void* dst, size_t bytes);
    cpq_status importState(const void* src, size_t bytes);
    cpq_status processCore(double* dIo, int64_t stride, int64_t T, unsigned stages, double* const* hostPlanar, float* const* hostF = nullptr);
    cpq_status processCoreImpl(double* dIo, int64_t stride, int64_t T, unsigned stages, double* const* hostPlanar, float* const* hostF,
                               bool deferSideStreams = false);
    // Time segments of a device-resident call with the dither branch (processCore): the serial shaper of segment s runs on the
    // side streams beside the transforms of segment s + 1 instead of trailing the whole call by one full-length chain.
    bool workStarted = false;            // the chunk loop of the last processCoreImpl call has begun (buffers may be modified)
    int forcedChunk = 0;                 // > 0: sequence chunk size of the first segment, kept for the others (chunk -> side stream map)
    int64_t segUniTotal = 0, segUniOffset = 0;   // injected uniforms: samples per channel of the whole call, first sample of the segment
    DevBuf<float> f32In, f32Out;        // device staging of float host buffers: 3 inbound / 2 outbound chunk slots
    PinnedBuf stageIn, stageOut;        // pinned staging rings for pageable FP64 host buffers (HostStager)
    cpq_status processDevice(double* dIo, int64_t stride, int64_t T, unsigned stages) { return processCore(dIo, stride, T, stages, nullptr); }
    cpq_status launchFwd(int log2P, const FwdArgs& a);
    cpq_status launchFwdLarge(int log2P, const FwdArgs& a);
    cpq_status launchInvLarge(int log2P, const InvArgs& a);
    DevBuf<double2> irScratchC;         // complex scratch for layer-2 IR partitions (P > 8192)
    cpq_status launchInv(int log2P, const InvArgs& a);
    cpq_status launchEq(EqArgs& a);
};

// ------------------------------------------------------------------------------------------------
// The >48 KB dynamic shared-memory opt-ins (cudaFuncAttributeMaxDynamicSharedMemorySize) are per device: Engine::init sets
// all of them after cudaSetDevice, for every handle, so a process may hold handles on several GPUs (setKernelAttributes).
template <int LOG2P>
static cudaError_t fwdLaunch(const FwdArgs& a, cudaStream_t s)
{
    using C = FftCfg<LOG2P>;
    const unsigned grid = (unsigned) ((a.totalFrames + C::FPC - 1) / C::FPC);
    const bool ir = a.halfOnly || a.applyScale || a.gain || a.tilt;   // prepare-time variant
    if (ir) fft_fwd_kernel<LOG2P, true><<<grid, C::THREADS, C::SMEM, s>>>(a);
    else fft_fwd_kernel<LOG2P, false><<<grid, C::THREADS, C::SMEM, s>>>(a);
    return cudaGetLastError();
}
template <int LOG2P>
static cudaError_t invLaunch(const InvArgs& a, cudaStream_t s)
{
    using C = FftCfg<LOG2P>;
    const unsigned grid = (unsigned) ((a.totalFrames + C::FPC - 1) / C::FPC);
    fft_inv_kernel<LOG2P><<<grid, C::THREADS, C::SMEM, s>>>(a);
    return cudaGetLastError();
}

// sixteen-points-per-thread kernels (cpq_fft16.cuh) for the streaming transforms at P = 512 / 4096
static const bool g_fft16 = [] { const char* e = getenv("CPQ_FFT16"); return !e || atoi(e) != 0; }();
template <int LOG2P>
static cudaError_t fwd16Launch(const FwdArgs& a, cudaStream_t s)
{
    using C = Fft16Cfg<LOG2P>;
    fft_fwd16_kernel<LOG2P><<<(unsigned) ((a.totalFrames + C::FPC - 1) / C::FPC), C::THREADS, C::SMEM, s>>>(a);
    return cudaGetLastError();
}
template <int LOG2P>
static cudaError_t inv16Launch(const InvArgs& a, cudaStream_t s)
{
    using C = Fft16Cfg<LOG2P>;
    fft_inv16_kernel<LOG2P><<<(unsigned) ((a.totalFrames + C::FPC - 1) / C::FPC), C::THREADS, C::SMEM, s>>>(a);
    return cudaGetLastError();
}

cpq_status Engine::launchFwd(int log2P, const FwdArgs& a)
{
    if (a.totalFrames <= 0) return CPQ_OK;
    cudaError_t e;
    const bool ir = a.halfOnly || a.applyScale || a.gain || a.tilt;
    if (g_fft16 && !ir && (log2P == 9 || log2P == 12))
    {
        e = log2P == 9 ? fwd16Launch<9>(a, stream) : fwd16Launch<12>(a, stream);
        ++launches;
        CPQ_CUDA(e);
        return CPQ_OK;
    }
    switch (log2P)
    {
        case 6: e = fwdLaunch<6>(a, stream); break;
        case 7: e = fwdLaunch<7>(a, stream); break;
        case 8: e = fwdLaunch<8>(a, stream); break;
        case 9: e = fwdLaunch<9>(a, stream); break;
        case 10: e = fwdLaunch<10>(a, stream); break;
        case 11: e = fwdLaunch<11>(a, stream); break;
        case 12: e = fwdLaunch<12>(a, stream); break;
        case 13: e = fwdLaunch<13>(a, stream); break;
        default: return launchFwdLarge(log2P, a);
    }
    ++launches;
    CPQ_CUDA(e);
    return CPQ_OK;
}
cpq_status Engine::launchInv(int log2P, const InvArgs& a)
{
    if (a.totalFrames <= 0) return CPQ_OK;
    cudaError_t e;
    if (g_fft16 && (log2P == 9 || log2P == 12))
    {
        e = log2P == 9 ? inv16Launch<9>(a, stream) : inv16Launch<12>(a, stream);
        ++launches;
        CPQ_CUDA(e);
        return CPQ_OK;
    }
    switch (log2P)
    {
        case 6: e = invLaunch<6>(a, stream); break;
        case 7: e = invLaunch<7>(a, stream); break;
        case 8: e = invLaunch<8>(a, stream); break;
        case 9: e = invLaunch<9>(a, stream); break;
        case 10: e = invLaunch<10>(a, stream); break;
        case 11: e = invLaunch<11>(a, stream); break;
        case 12: e = invLaunch<12>(a, stream); break;
        case 13: e = invLaunch<13>(a, stream); break;
        default: return launchInvLarge(log2P, a);
    }
    ++launches;
    CPQ_CUDA(e);
    return CPQ_OK;
}

// ---- P > 8192: one kernel per radix pass, ping-pong between two global buffers (cpq_fft_large.cuh) ----
template <int SIGN>
static cudaError_t largePasses(const LargeFftArgs& la, double2*& cur, double2*& other, int log2P, cudaStream_t s, int64_t& launches)
{
    const int P = 1 << log2P;
    int Ns = 1;
    const int r0 = 1 << (log2P % 3);
    auto grid = [&](int64_t n) { return (unsigned) ((n + 255) / 256); };
    if (r0 == 2) { gfft_pass_kernel<2, SIGN><<<grid(la.totalFrames * (P / 2)), 256, 0, s>>>(la, cur, other, Ns); Ns *= 2; std::swap(cur, other); ++launches; }
    else if (r0 == 4) { gfft_pass_kernel<4, SIGN><<<grid(la.totalFrames * (P / 4)), 256, 0, s>>>(la, cur, other, Ns); Ns *= 4; std::swap(cur, other); ++launches; }
    for (int p = 0; p < log2P / 3; ++p)
    {
        gfft_pass_kernel<8, SIGN><<<grid(la.totalFrames * (P / 8)), 256, 0, s>>>(la, cur, other, Ns);
        Ns *= 8;
        std::swap(cur, other);
        ++launches;
    }
    return cudaGetLastError();
}

cpq_status Engine::launchFwdLarge(int log2P, const FwdArgs& a)
{
    if (log2P > 16 || !a.scratch)
    {
        setError("partition size > 65536 (FFT > 131072) is not supported");
        return CPQ_ERR_UNSUPPORTED;
    }
    LargeFftArgs la {};
    la.P = 1 << log2P;
    la.totalFrames = a.totalFrames;
    la.framesPerSeq = a.framesPerSeq;
    la.src = a.src; la.srcStride = a.srcStride; la.frameStart0 = a.frameStart0; la.lo = a.lo; la.hi = a.hi; la.halfOnly = a.halfOnly;
    la.histEnd = a.histEnd; la.histStride = a.histStride;
    la.rowPitchFrames = a.outFramesPerSeq; la.rowOffset = a.outFrameOffset;
    la.tw = a.tw; la.scale = a.scale; la.applyScale = a.applyScale; la.gain = a.gain; la.tilt = a.tilt;
    static const int g2Env = [] { const char* e = getenv("CPQ_GFFT2"); return e ? atoi(e) : 1; }();   // 0: one kernel per radix pass (round 1)
    if (g2Env && log2P >= 14)
    {
        // four-step transform: column pass straight from the signal into scratch, row pass into out, split in place
        la.a = a.out;
        const int log2N1 = log2P - 8;
        const unsigned gc = (unsigned) (la.totalFrames * (kG2N2 / kG2Cols)), gr = (unsigned) (la.totalFrames * ((1 << log2N1) / kG2Rows));
        const size_t sc = ((size_t) kG2Cols * ((1 << log2N1) + 1) + (1 << log2N1)) * sizeof(double2), sr = ((size_t) kG2Rows * (kG2N2 + 1) + kG2N2) * sizeof(double2);
        switch (log2N1)
        {
            case 6: gfft2_cols_kernel<6, -1, true><<<gc, kG2Cols * 8, sc, stream>>>(la, nullptr, a.scratch); gfft2_rows_kernel<6, -1, false><<<gr, kG2Rows * 32, sr, stream>>>(la, a.scratch, a.out); break;
            case 7: gfft2_cols_kernel<7, -1, true><<<gc, kG2Cols * 16, sc, stream>>>(la, nullptr, a.scratch); gfft2_rows_kernel<7, -1, false><<<gr, kG2Rows * 32, sr, stream>>>(la, a.scratch, a.out); break;
            default: gfft2_cols_kernel<8, -1, true><<<gc, kG2Cols * 32, sc, stream>>>(la, nullptr, a.scratch); gfft2_rows_kernel<8, -1, false><<<gr, kG2Rows * 32, sr, stream>>>(la, a.scratch, a.out); break;
        }
        launches += 2;
        CPQ_CUDA(cudaGetLastError());
        const int64_t h2 = la.totalFrames * (la.P / 2 + 1);
        gfft_split_fwd_kernel<<<(unsigned) ((h2 + 255) / 256), 256, 0, stream>>>(la, a.out);
        ++launches;
        CPQ_CUDA(cudaGetLastError());
        return CPQ_OK;
    }
    const int nPasses = (log2P % 3 ? 1 : 0) + log2P / 3;
    // start in the buffer that makes the last pass land in a.out
    double2* cur = (nPasses % 2 == 0) ? a.out : a.scratch;
    double2* other = (nPasses % 2 == 0) ? a.scratch : a.out;
    la.a = cur;
    const int64_t n = la.totalFrames * la.P;
    gfft_load_fwd_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(la);
    ++launches;
    CPQ_CUDA(cudaGetLastError());
    CPQ_CUDA(largePasses<-1>(la, cur, other, log2P, stream, launches));
    const int64_t h = la.totalFrames * (la.P / 2 + 1);
    gfft_split_fwd_kernel<<<(unsigned) ((h + 255) / 256), 256, 0, stream>>>(la, cur);   // cur == a.out
    ++launches;
    CPQ_CUDA(cudaGetLastError());
    return CPQ_OK;
}

cpq_status Engine::launchInvLarge(int log2P, const InvArgs& a)
{
    if (log2P > 16 || !a.scratch)
    {
        setError("partition size > 65536 (FFT > 131072) is not supported");
        return CPQ_ERR_UNSUPPORTED;
    }
    LargeFftArgs la {};
    la.P = 1 << log2P;
    la.totalFrames = a.totalFrames;
    la.framesPerSeq = a.framesOut;
    la.rowPitchFrames = a.framesPerSeq; la.rowOffset = 0;
    la.dst = a.out; la.dstStride = a.outStride;
    la.tw = a.tw;
    double2* cur = const_cast<double2*>(a.in);
    double2* other = a.scratch;
    const int64_t h = la.totalFrames * (la.P / 2 + 1);
    gfft_pre_inv_kernel<<<(unsigned) ((h + 255) / 256), 256, 0, stream>>>(la, cur);
    ++launches;
    CPQ_CUDA(cudaGetLastError());
    static const int g2Env = [] { const char* e = getenv("CPQ_GFFT2"); return e ? atoi(e) : 1; }();
    if (g2Env && log2P >= 14)
    {
        // four-step inverse: column pass in -> scratch, row pass scratch -> the kept half of every frame as real samples
        const int log2N1 = log2P - 8;
        const unsigned gc = (unsigned) (la.totalFrames * (kG2N2 / kG2Cols)), gr = (unsigned) (la.totalFrames * ((1 << log2N1) / kG2Rows));
        const size_t sc = ((size_t) kG2Cols * ((1 << log2N1) + 1) + (1 << log2N1)) * sizeof(double2), sr = ((size_t) kG2Rows * (kG2N2 + 1) + kG2N2) * sizeof(double2);
        switch (log2N1)
        {
            case 6: gfft2_cols_kernel<6, +1, false><<<gc, kG2Cols * 8, sc, stream>>>(la, cur, other); gfft2_rows_kernel<6, +1, true><<<gr, kG2Rows * 32, sr, stream>>>(la, other, nullptr); break;
            case 7: gfft2_cols_kernel<7, +1, false><<<gc, kG2Cols * 16, sc, stream>>>(la, cur, other); gfft2_rows_kernel<7, +1, true><<<gr, kG2Rows * 32, sr, stream>>>(la, other, nullptr); break;
            default: gfft2_cols_kernel<8, +1, false><<<gc, kG2Cols * 32, sc, stream>>>(la, cur, other); gfft2_rows_kernel<8, +1, true><<<gr, kG2Rows * 32, sr, stream>>>(la, other, nullptr); break;
        }
        launches += 2;
        CPQ_CUDA(cudaGetLastError());
        return CPQ_OK;
    }
    CPQ_CUDA(largePasses<+1>(la, cur, other, log2P, stream, launches));
    const int64_t n = la.totalFrames * (la.P / 2);
    gfft_store_inv_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(la, cur);
    ++launches;
    CPQ_CUDA(cudaGetLastError());
    return CPQ_OK;
}

constexpr size_t kMaxDynSmem = 227 * 1024;   // opt-in limit per CTA on sm_100

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encodeTiledFn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// rank-3 FP64 tensor [d2][d1][d0 doubles] with a box of [1][box1][64 doubles]; out-of-bounds elements read as zero
// rank-3 map [d2 sequences][d1 rows][d0 doubles]; pitchRows = rows between two sequences when that is more than the d1 rows the
// map may touch (a window into a larger per-sequence buffer)
static bool encodeSpectraMap(MacTensorMap& out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box1, uint64_t pitchRows = 0)
{
    EncodeTiledFn fn = encodeTiledFn();
    if (!fn) return false;
    static_assert(sizeof(CUtensorMap) == sizeof(MacTensorMap), "tensor map size");
    CUtensorMap tm;
    const cuuint64_t gdim[3] = { d0, d1, d2 };
    const cuuint64_t gstride[2] = { d0 * sizeof(double), d0 * (pitchRows ? pitchRows : d1) * sizeof(double) };
    const cuuint32_t box[3] = { 64, box1, 1 };
    const cuuint32_t estride[3] = { 1, 1, 1 };
    if (fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    std::memcpy(out.bytes, &tm, sizeof(tm));
    return true;
}

template <int LOG2P>
static cudaError_t fftAttrs()
{
    cudaError_t e = cudaFuncSetAttribute(fft_fwd_kernel<LOG2P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FftCfg<LOG2P>::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fft_fwd_kernel<LOG2P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FftCfg<LOG2P>::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fft_inv_kernel<LOG2P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FftCfg<LOG2P>::SMEM);
    return e;
}
template <int LOG2P>
static cudaError_t fft16Attrs()
{
    cudaError_t e = cudaFuncSetAttribute(fft_fwd16_kernel<LOG2P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Fft16Cfg<LOG2P>::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fft_inv16_kernel<LOG2P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Fft16Cfg<LOG2P>::SMEM);
    return e;
}

// Per device (the attribute lives in the device's context), hence per handle and unconditional.
cpq_status Engine::setKernelAttributes()
{
    CPQ_CUDA(fftAttrs<6>()); CPQ_CUDA(fftAttrs<7>()); CPQ_CUDA(fftAttrs<8>()); CPQ_CUDA(fftAttrs<9>());
    CPQ_CUDA(fftAttrs<10>()); CPQ_CUDA(fftAttrs<11>()); CPQ_CUDA(fftAttrs<12>()); CPQ_CUDA(fftAttrs<13>());
    CPQ_CUDA(fft16Attrs<9>()); CPQ_CUDA(fft16Attrs<12>());
    CPQ_CUDA(cudaFuncSetAttribute(eq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kEqSmemBytes));
    CPQ_CUDA(cudaFuncSetAttribute(eq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kEqSmemBytesPost));
    CPQ_CUDA(cudaFuncSetAttribute(eq_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kEqSmemBytesPost));
    CPQ_CUDA(cudaFuncSetAttribute(eq_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kEqSmemBytesPost));
    CPQ_CUDA(cudaFuncSetAttribute(eq_kernel<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kEqSmemBytes));
    CPQ_CUDA(cudaFuncSetAttribute(eq_kernel<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kEqSmemBytesPost));
    CPQ_CUDA(cudaFuncSetAttribute(eq_kernel<true, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kEqSmemBytesPost));
    CPQ_CUDA(cudaFuncSetAttribute(eq_kernel<true, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kEqSmemBytesPost));
    CPQ_CUDA(cudaFuncSetAttribute(mac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem));
    CPQ_CUDA(cudaFuncSetAttribute(mac_tma_kernel<kMacGroups>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem));
    CPQ_CUDA(cudaFuncSetAttribute(mac_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem));
    CPQ_CUDA(cudaFuncSetAttribute(dither_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kDitherSmemBytes));
    {
        const int g2 = (int) ((kG2Cols * 257 + 256) * sizeof(double2));   // column pass at N1 = 256; N1 = 128 needs 68 KB, N1 = 64 fits the default
        CPQ_CUDA(cudaFuncSetAttribute(gfft2_cols_kernel<8, -1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g2));
        CPQ_CUDA(cudaFuncSetAttribute(gfft2_cols_kernel<8, +1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g2));
        CPQ_CUDA(cudaFuncSetAttribute(gfft2_cols_kernel<7, -1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g2));
        CPQ_CUDA(cudaFuncSetAttribute(gfft2_cols_kernel<7, +1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g2));
    }
    return CPQ_OK;
}

cpq_status Engine::init(const cpq_config* c)
{
    cfg = *c;
    if (cfg.n_streams <= 0 || cfg.n_channels < 1 || cfg.n_channels > 2 || cfg.sample_rate <= 0.0 || cfg.max_samples <= 0)
    {
        setError("invalid config");
        return CPQ_ERR_INVALID;
    }
    // Power-of-two host blocks are the reference's regular case (L0 partition == block, zero latency).  Other host blocks run
    // like the application does: SetImpulse gets the block rounded up to a power of two (knownBlockSize) while Add/Get are
    // called with the host block (preferredCallSize), LoaderThread.cpp:230,239-245 -- L0 then goes through its output ring.
    if (cfg.block_size < 64 || cfg.block_size > 8192)
    {
        setError("block_size must be in 64..8192");
        return CPQ_ERR_UNSUPPORTED;
    }
    if (cfg.max_samples % cfg.block_size != 0)
    {
        setError("max_samples must be a multiple of block_size");
        return CPQ_ERR_INVALID;
    }
    nSeq = cfg.n_streams * cfg.n_channels;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
    {
        cudaGetLastError();
        setError("no CUDA device: convopeq_b200 has no CPU fallback");
        return CPQ_ERR_CUDA;
    }
    if (cfg.device < 0 || cfg.device >= count)
    {
        setError("device ordinal out of range");
        return CPQ_ERR_INVALID;
    }
    CPQ_CUDA(cudaSetDevice(cfg.device));
    cudaDeviceProp prop {};
    CPQ_CUDA(cudaGetDeviceProperties(&prop, cfg.device));
    if (prop.major < 10)
    {
        setError("device is not sm_100 class (B200 required)");
        return CPQ_ERR_CUDA;
    }
    if (cfg.workspace_bytes == 0)
    {
        // default: a quarter of the device memory that is free now, within 4..16 GiB (measured on cfg4: 2 GiB 78.0 ms,
        // 4 GiB 74.7 ms, 8 GiB 73.4 ms, 16 GiB 73.2 ms per step -- fewer, larger launches have shorter tails)
        size_t freeB = 0, totalB = 0;
        cfg.workspace_bytes = (size_t) 4 << 30;
        if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess)
            cfg.workspace_bytes = std::min<size_t>((size_t) 16 << 30, std::max<size_t>((size_t) 4 << 30, freeB / 4));
    }
    {
        cpq_status st = setKernelAttributes();
        if (st != CPQ_OK) return st;
    }
    CPQ_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CPQ_CUDA(cudaStreamCreateWithFlags(&sIn, cudaStreamNonBlocking));
    CPQ_CUDA(cudaStreamCreateWithFlags(&sOut, cudaStreamNonBlocking));
    {
        // highest priority: when an SM frees resources the waiting dither CTA (one warp) is placed before the next CTA of the
        // running transform, so the serial stage starts as early as it can and runs beside the rest of the call
        int prLow = 0, prHigh = 0;
        CPQ_CUDA(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
        for (auto& d : sDither) CPQ_CUDA(cudaStreamCreateWithPriority(&d, cudaStreamNonBlocking, prHigh));
    }
    for (auto& e : ev) CPQ_CUDA(cudaEventCreate(&e));
    nH = cfg.shared_ir ? cfg.n_channels : nSeq;
    haveImpulse.assign((size_t) nH, 0);
    eqSets.assign((size_t) (cfg.shared_eq ? 1 : cfg.n_streams), EqSet {});
    CPQ_CUDA(ticketFault.ensure(2));
    CPQ_CUDA(cudaMemsetAsync(ticketFault.p, 0, 2 * sizeof(unsigned), stream));
    CPQ_CUDA(ditherZ.ensure((size_t) nSeq * 12));
    CPQ_CUDA(cudaMemsetAsync(ditherZ.p, 0, (size_t) nSeq * 12 * sizeof(double), stream));
    CPQ_CUDA(stateOut.ensure((size_t) nSeq * CPQ_NUM_BANDS * 2));
    CPQ_CUDA(cudaMemsetAsync(stateOut.p, 0, (size_t) nSeq * CPQ_NUM_BANDS * 2 * sizeof(double), stream));
    CPQ_CUDA(cudaStreamSynchronize(stream));
    return CPQ_OK;
}

cpq_status Engine::ensureTwiddles(int li)
{
    LayerDev& L = layer[li];
    const int P = plan.layers[li].partSize;
    if (L.tw.p && L.tw.n == 3 * ((size_t) P + 1)) return CPQ_OK;
    // three tables of P+1 entries: W = exp(-2 pi i t / 2P); -i/2 W (forward split pass); i conj(W) / 2P (inverse pre-pass)
    std::vector<double2> tw(3 * ((size_t) P + 1));
    const long double twoPiOverN = 2.0L * 3.141592653589793238462643383279502884L / (long double) (2 * P);
    for (int t = 0; t <= P; ++t)
    {
        // octant symmetry keeps the table exact at the special angles
        long double c, s;
        if (t == 0) { c = 1.0L; s = 0.0L; }
        else if (t == P) { c = -1.0L; s = 0.0L; }
        else if (2 * t == P) { c = 0.0L; s = 1.0L; }
        else { c = cosl(twoPiOverN * t); s = sinl(twoPiOverN * t); }
        const double2 w = make_double2((double) c, (double) -s);
        tw[(size_t) t] = w;
        tw[(size_t) (P + 1) + t] = make_double2(0.5 * w.y, -0.5 * w.x);
        tw[2 * (size_t) (P + 1) + t] = make_double2(w.y / (double) (2 * P), w.x / (double) (2 * P));
    }
    CPQ_CUDA(L.tw.ensure(tw.size()));
    CPQ_CUDA(cudaMemcpyAsync(L.tw.p, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
    // compact per-pass tables {W^k, W^2k, W^4k}, W = exp(-2 pi i / (8 Ns)), for the radix-8 passes (FftCfg::passOffset)
    std::vector<double2> ptw;
    {
        const int log2P = ilog2(P);
        int ns = 1 << (log2P % 3);
        const long double twoPi = 2.0L * 3.141592653589793238462643383279502884L;
        for (int p = 0; p < log2P / 3; ++p, ns *= 8)
        {
            if (ns <= 1) continue;
            for (int c = 1; c <= 4; c *= 2)
                for (int k = 0; k < ns; ++k)
                {
                    const long double ang = twoPi * (long double) (c * k) / (long double) (8 * ns);
                    ptw.push_back(make_double2((double) cosl(ang), (double) -sinl(ang)));
                }
        }
        if (ptw.empty()) ptw.push_back(make_double2(1.0, 0.0));
    }
    CPQ_CUDA(L.ptw.ensure(ptw.size()));
    CPQ_CUDA(cudaMemcpyAsync(L.ptw.p, ptw.data(), ptw.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
    CPQ_CUDA(cudaStreamSynchronize(stream));
    return CPQ_OK;
}

cpq_status Engine::setImpulse(int stream_, int ch, const double* ir, int len, double scale, const cpq_filter_spec* spec)
{
    if (!ir || len <= 0 || ch < 0 || ch >= cfg.n_channels)
    {
        setError("set_impulse: bad argument");
        return CPQ_ERR_INVALID;
    }
    if (cfg.shared_ir ? (stream_ != -1 && stream_ != 0) : (stream_ < 0 || stream_ >= cfg.n_streams))
    {
        setError("set_impulse: stream out of range");
        return CPQ_ERR_INVALID;
    }
    ConvPlan p;
    int knownBlock = 64;
    while (knownBlock < cfg.block_size) knownBlock <<= 1;   // juce::nextPowerOfTwo(host block)
    const bool planOk = makeConvPlan(len, knownBlock, spec, p, cfg.uniform_partitions != 0);
    p.callSize = cfg.block_size;
    if (!planOk)
    {
        setError("set_impulse: SetImpulse would reject these parameters");
        return CPQ_ERR_INVALID;
    }
    if (cfg.uniform_partitions && p.layers[0].numPartsIR > 160)
    {
        setError("set_impulse: the uniform-partition extension holds at most 160 partitions (the MAC kernel's shared-memory tile)");
        return CPQ_ERR_UNSUPPORTED;
    }
    for (int li = 0; li < p.numLayers; ++li)
        if (p.layers[li].partSize > 65536)
        {
            setError("set_impulse: layer partition > 65536 (FFT > 131072) is not supported");
            return CPQ_ERR_UNSUPPORTED;
        }
    if (planSet && !plan.sameGeometry(p))
    {
        setError("set_impulse: every stream-channel of a handle must share the layer geometry");
        return CPQ_ERR_GEOMETRY;
    }
    CPQ_CUDA(cudaSetDevice(cfg.device));
    const bool first = !planSet;
    if (first)
    {
        plan = p;
        planSet = true;
        gplanCallbacks = -1;
        gplanHostCallbacks = -1;
        for (int li = 0; li < plan.numLayers; ++li)
        {
            const LayerPlan& l = plan.layers[li];
            cpq_status st = ensureTwiddles(li);
            if (st != CPQ_OK) return st;
            CPQ_CUDA(layer[li].H.ensure((size_t) nH * l.numPartsIR * l.partSize));
            CPQ_CUDA(cudaMemsetAsync(layer[li].H.p, 0, (size_t) nH * l.numPartsIR * l.partSize * sizeof(double2), stream));
        }
    }
    // Per-bin gains depend on the FilterSpec, which may differ between channels only in ways that keep the
    // geometry; rebuild them for this call.
    const int row = hRowOf(stream_ < 0 ? 0 : stream_, ch);
    CPQ_CUDA(irScratch.ensure((size_t) len + 2));
    CPQ_CUDA(cudaMemcpyAsync(irScratch.p, ir, (size_t) len * sizeof(double), cudaMemcpyHostToDevice, stream));
    if (directHead)
    {
        // m_directTapCount = min(irLen, min(nextPow2(max(block, 64)), 32)) = min(irLen, 32); m_directIRRev[i] = impulse[taps-1-i] * scale
        // (:689-718); the partitions are built from the impulse with those taps zeroed (:730-731)
        const int taps = std::min(len, 32);
        double rev[32] = {};
        for (int i = 0; i < taps; ++i) rev[32 - taps + i] = ir[taps - 1 - i] * scale;
        CPQ_CUDA(directTaps.ensure((size_t) nH * 32));
        CPQ_CUDA(cudaMemcpyAsync(directTaps.p + (size_t) row * 32, rev, sizeof(rev), cudaMemcpyHostToDevice, stream));
        CPQ_CUDA(cudaMemsetAsync(irScratch.p, 0, (size_t) taps * sizeof(double), stream));
        CPQ_CUDA(cudaStreamSynchronize(stream));   // rev is on the stack
    }
    for (int li = 0; li < p.numLayers; ++li)
    {
        const LayerPlan& l = p.layers[li];
        LayerDev& L = layer[li];
        std::vector<double> g, t;
        const double* dGain = nullptr;
        const double* dTilt = nullptr;
        if (p.hasSpec)
        {
            spectrumGain(p, li, g);
            CPQ_CUDA(L.gain.ensure(g.size()));
            CPQ_CUDA(cudaMemcpyAsync(L.gain.p, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
            dGain = L.gain.p;
        }
        if (tiltGain(p, li, t))
        {
            CPQ_CUDA(L.tilt.ensure(t.size()));
            CPQ_CUDA(cudaMemcpyAsync(L.tilt.p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
            dTilt = L.tilt.p;
        }
        FwdArgs a {};
        a.src = irScratch.p;
        a.srcStride = 0;
        a.frameStart0 = l.irOffset;
        a.lo = l.irOffset;
        a.hi = (int64_t) l.irOffset + l.irLen;
        a.halfOnly = 1;
        a.framesPerSeq = l.numPartsIR;
        a.totalFrames = l.numPartsIR;
        a.out = L.H.p + (size_t) row * l.numPartsIR * l.partSize;
        a.outFramesPerSeq = l.numPartsIR;
        a.outFrameOffset = 0;
        a.tw = L.tw.p;
        a.ptw = L.ptw.p;
        a.scale = scale;
        a.applyScale = std::fabs(scale - 1.0) > 1e-12 ? 1 : 0;
        a.gain = dGain;
        a.tilt = dTilt;
        if (l.partSize > 8192)
        {
            CPQ_CUDA(irScratchC.ensure((size_t) l.numPartsIR * l.partSize));
            a.scratch = irScratchC.p;
        }
        cpq_status st = launchFwd(ilog2(l.partSize), a);
        if (st != CPQ_OK) return st;
        // the host vectors g/t must outlive the async copies
        CPQ_CUDA(cudaStreamSynchronize(stream));
    }
    haveImpulse[(size_t) row] = 1;
    return CPQ_OK;
}

static void matmul2(const long double* a, const long double* b, long double* c)
{
    long double r[4] = { a[0] * b[0] + a[1] * b[2], a[0] * b[1] + a[1] * b[3], a[2] * b[0] + a[3] * b[2], a[2] * b[1] + a[3] * b[3] };
    for (int i = 0; i < 4; ++i) c[i] = r[i];
}

static void buildScanTables(const long double A[4], const long double b[2], double* out);

static void buildBandConstants(const cpq_svf_coeffs& c, double* out /* kEqcStride */)
{
    std::memset(out, 0, sizeof(double) * kEqcStride);
    out[0] = c.a1; out[1] = c.a2; out[2] = c.a3; out[3] = c.m0; out[4] = c.m1; out[5] = c.m2;
    // s' = A s + b v0   (EQProcessor.Processing.cpp:148-153 in affine form)
    const long double A[4] = { 2.0L * c.a1 - 1.0L, -2.0L * c.a2, 2.0L * c.a2, 1.0L - 2.0L * c.a3 };
    const long double b[2] = { 2.0L * c.a2, 2.0L * c.a3 };
    {
        // The fast pass relies on ||A^n|| <= 1 (true for every SVF the reference designs); arbitrary user
        // coefficients that break it are always replayed with the exact per-sample semantics.
        const double p = (double) (A[0] * A[0] + A[2] * A[2]), q = (double) (A[0] * A[1] + A[2] * A[3]);
        const double r = (double) (A[1] * A[1] + A[3] * A[3]);
        const double lmax = 0.5 * (p + r) + std::sqrt(0.25 * (p - r) * (p - r) + q * q);   // largest eigenvalue of A^T A
        out[6] = (std::isfinite(lmax) && lmax <= 1.0 + 1e-9) ? 0.0 : 1.0;
        // TPT consistency (calcSVFCoeffs: a2 = g a1, a3 = g a2): lets pass 2 use v2 = ic2 + g v1.  g is recovered from
        // the coefficients themselves so that raw cpq_set_eq input is classified the same way as designed bands.
        const double g = (c.a1 != 0.0) ? c.a2 / c.a1 : 0.0;
        const bool tpt = c.a1 != 0.0 && std::isfinite(g) && std::fabs(c.a3 - g * c.a2) <= 8.0 * 2.220446049250313e-16 * std::fabs(c.a3);
        out[7] = !tpt ? 0.0 : ((c.m0 == 1.0 && c.m2 == 0.0) ? 1.0 : 2.0);   // 1: Peaking pattern, out = v0 + m1 v1
        out[8] = g;
        out[9] = 2.0 * g;
        out[10] = (double) (2.0L * ((long double) c.a1 - (long double) c.a3));   // rotated-state recurrence of the Peaking fast path
        out[11] = (double) (2.0L * ((long double) c.a1 + (long double) c.a3));
    }
    buildScanTables(A, b, out);
}

// Tables of the blocked scan of s' = A s + b x (cpq_eq.cuh): zero-state weights and the powers of A the warp scan,
// the lane lookup and the segment link need; long double so that A^1024 is correctly rounded to double.
static void buildScanTables(const long double A[4], const long double b[2], double* out)
{
    // w[j] = A^(L-1-j) b
    long double v[2] = { b[0], b[1] };
    for (int j = kEqL - 1; j >= 0; --j)
    {
        out[kEqcW + 2 * j] = (double) v[0];
        out[kEqcW + 2 * j + 1] = (double) v[1];
        const long double n0 = A[0] * v[0] + A[1] * v[1], n1 = A[2] * v[0] + A[3] * v[1];
        v[0] = n0; v[1] = n1;
    }
    long double AL[4] = { 1, 0, 0, 1 };
    for (int i = 0; i < kEqL; ++i)
    {
        if (i == kEqL / 2)
            for (int k = 0; k < 4; ++k) out[kEqcMh + k] = (double) AL[k];   // A^(L/2): half-block step of the two-chain pass 2
        matmul2(AL, A, AL);
    }
    // Plo[j] = A^(L j), j < 8;  Phi[j] = A^(8L j), j < 4
    {
        long double M[4] = { 1, 0, 0, 1 };
        for (int j = 0; j < 8; ++j)
        {
            for (int i = 0; i < 4; ++i) out[kEqcPlo + 4 * j + i] = (double) M[i];
            matmul2(M, AL, M);
        }
        long double A8L[4] = { M[0], M[1], M[2], M[3] };
        long double H[4] = { 1, 0, 0, 1 };
        for (int j = 0; j < 4; ++j)
        {
            for (int i = 0; i < 4; ++i) out[kEqcPhi + 4 * j + i] = (double) H[i];
            matmul2(H, A8L, H);
        }
    }
    long double M[4] = { AL[0], AL[1], AL[2], AL[3] };
    for (int d = 0; d < 5; ++d)
    {
        for (int i = 0; i < 4; ++i) out[kEqcMs + 4 * d + i] = (double) M[i];
        matmul2(M, M, M);
    }
    // M is now A^(32 L): one warp segment
    for (int i = 0; i < 4; ++i) out[kEqcMw + i] = (double) M[i];
    // Pw[j] = A^(32 L j), j = 0..8: whole warp segments, for the look-back composition of the segment start states
    long double W[4] = { 1, 0, 0, 1 };
    for (int j = 0; j <= 8; ++j)
    {
        for (int i = 0; i < 4; ++i) out[kEqcPw + 4 * j + i] = (double) W[i];
        matmul2(W, M, W);
    }
}

// ---- linear output stages: OutputFilter (OutputFilter.cpp:28-112) and the output DC blocker (UltraHighRateDCBlocker.h:60-90) ----
struct BiquadC { double b0 = 1.0, b1 = 0.0, b2 = 0.0, a1 = 0.0, a2 = 0.0; };
static BiquadC makeLPF(double fc, double Q, double fs)   // OutputFilter::makeLPF, :28-48 (RBJ, a0-normalised)
{
    BiquadC c;
    const double nyq = fs * 0.4999;
    if (fc >= nyq || Q <= 0.0 || fs <= 0.0) return c;
    const double w0 = 2.0 * 3.14159265358979323846 * fc / fs;
    const double sn = std::sin(w0), cs = std::cos(w0);
    const double alpha = sn / (2.0 * Q);
    const double a0inv = 1.0 / (1.0 + alpha);
    c.b0 = (1.0 - cs) * 0.5 * a0inv;
    c.b1 = (1.0 - cs) * a0inv;
    c.b2 = (1.0 - cs) * 0.5 * a0inv;
    c.a1 = (-2.0 * cs) * a0inv;
    c.a2 = (1.0 - alpha) * a0inv;
    return c;
}
static BiquadC makeHPF(double fc, double Q, double fs)   // OutputFilter::makeHPF, :50-70
{
    BiquadC c;
    const double nyq = fs * 0.4999;
    if (fc <= 0.0 || fc >= nyq || Q <= 0.0 || fs <= 0.0) return c;
    const double w0 = 2.0 * 3.14159265358979323846 * fc / fs;
    const double sn = std::sin(w0), cs = std::cos(w0);
    const double alpha = sn / (2.0 * Q);
    const double a0inv = 1.0 / (1.0 + alpha);
    c.b0 = (1.0 + cs) * 0.5 * a0inv;
    c.b1 = -(1.0 + cs) * a0inv;
    c.b2 = (1.0 + cs) * 0.5 * a0inv;
    c.a1 = (-2.0 * cs) * a0inv;
    c.a2 = (1.0 - alpha) * a0inv;
    return c;
}
// the three cascaded stages OutputFilter::process runs, in order (OutputFilter.cpp:160-166 / :290-296; prepare :80-112)
static void outputFilterStages(double sr, int convIsLast, int hcMode, int lcMode, int lpMode, BiquadC out[3])
{
    const double fcHc = (sr <= 48000.0) ? 19000.0 : 22000.0;
    const double fcLp = (sr <= 48000.0) ? 19000.0 : 24000.0;
    if (convIsLast)
    {
        out[0] = lcMode == 1 ? makeHPF(15.0, 0.5, sr) : makeHPF(18.0, 0.70711, sr);
        if (hcMode == 0) { out[1] = makeLPF(fcHc, 0.54120, sr); out[2] = makeLPF(fcHc, 1.30656, sr); }
        else if (hcMode == 2) { out[1] = makeLPF(fcHc, 0.5, sr); out[2] = BiquadC {}; }
        else { out[1] = makeLPF(fcHc, 0.70711, sr); out[2] = makeLPF(fcHc, 0.70711, sr); }
    }
    else
    {
        out[0] = makeHPF(20.0, 0.70711, sr);
        const double q = lpMode == 0 ? 1.0 : (lpMode == 2 ? 0.5 : 0.70711);
        out[1] = makeLPF(fcLp, q, sr);
        out[2] = makeLPF(fcLp, q, sr);
    }
}
static void dcBlockerAlphas(double sr, double cutoffHz, double alpha[2])   // UltraHighRateDCBlocker::init, :60-90
{
    alpha[0] = alpha[1] = 1.0e-6;
    if (!std::isfinite(sr) || sr <= 0.0 || !std::isfinite(cutoffHz) || cutoffHz <= 0.0) return;
    const double ratios[2] = { 1.0 - 0.1, 1.0 + 0.1 };
    for (int i = 0; i < 2; ++i)
    {
        const double omega = 2.0 * 3.14159265358979323846 * (cutoffHz * ratios[i]) / sr;
        double a = -std::expm1(-omega);
        if (!std::isfinite(a) || a <= 0.0 || a >= 1.0) a = 1.0e-6;
        alpha[i] = a;
    }
}
static void buildBiquadConstants(const BiquadC& c, double* out)
{
    std::memset(out, 0, sizeof(double) * kEqcStride);
    out[0] = c.b0; out[1] = c.b1; out[2] = c.b2; out[3] = c.a1; out[4] = c.a2;
    out[7] = 3.0;
    // DF2T: y = b0 x + w1; w1' = b1 x - a1 y + w2; w2' = b2 x - a2 y  ->  s' = A s + b x
    const long double A[4] = { -(long double) c.a1, 1.0L, -(long double) c.a2, 0.0L };
    const long double b[2] = { (long double) c.b1 - (long double) c.a1 * c.b0, (long double) c.b2 - (long double) c.a2 * c.b0 };
    buildScanTables(A, b, out);
}
static void buildDcConstants(const double alpha[2], double* out)
{
    std::memset(out, 0, sizeof(double) * kEqcStride);
    out[0] = alpha[0]; out[1] = alpha[1];
    out[7] = 4.0;
    // s0' = (1-a0) s0 + a0 x;  x1 = x - s0';  s1' = (1-a1) s1 + a1 x1
    const long double a0 = alpha[0], a1 = alpha[1];
    const long double A[4] = { 1.0L - a0, 0.0L, -a1 * (1.0L - a0), 1.0L - a1 };
    const long double b[2] = { a0, a1 * (1.0L - a0) };
    buildScanTables(A, b, out);
}

cpq_status Engine::uploadPost()
{
    if (!postDirty) return CPQ_OK;
    std::vector<double> host((size_t) kEqPostStages * kEqcStride, 0.0);
    BiquadC st[3];
    outputFilterStages(cfg.sample_rate, outCfg.convIsLast, outCfg.hc, outCfg.lc, outCfg.lp, st);
    postIdentity = 0;
    for (int i = 0; i < 3; ++i)
    {
        buildBiquadConstants(st[i], host.data() + (size_t) i * kEqcStride);
        if (st[i].b0 == 1.0 && st[i].b1 == 0.0 && st[i].b2 == 0.0 && st[i].a1 == 0.0 && st[i].a2 == 0.0) postIdentity |= 1u << i;
    }
    double alpha[2];
    dcBlockerAlphas(cfg.sample_rate, outCfg.dcCutoff, alpha);
    buildDcConstants(alpha, host.data() + (size_t) 3 * kEqcStride);
    CPQ_CUDA(postc.ensure(host.size()));
    CPQ_CUDA(cudaMemcpyAsync(postc.p, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    CPQ_CUDA(cudaStreamSynchronize(stream));
    postDirty = false;
    return CPQ_OK;
}

cpq_status Engine::uploadEq(int64_t nCallbacks)
{
    const size_t nSets = eqSets.size();
    if (eqDirty)
    {
        std::vector<double> host(nSets * CPQ_NUM_BANDS * kEqcStride);
        std::vector<double> sat(nSets), gc(nSets);
        for (size_t s = 0; s < nSets; ++s)
        {
            if (!eqSets[s].set)
            {
                setError("process: EQ stage requested but cpq_set_eq was not called for every stream");
                return CPQ_ERR_NOT_READY;
            }
            for (int b = 0; b < CPQ_NUM_BANDS; ++b)
                buildBandConstants(eqSets[s].coeffs[b], host.data() + (s * CPQ_NUM_BANDS + b) * kEqcStride);
            sat[s] = eqSets[s].saturation;
            gc[s] = eqSets[s].totalGain;
        }
        std::vector<unsigned> mask((size_t) nSeq), scalar((size_t) nSeq);
        std::vector<int> sos((size_t) nSeq);
        anyPar = anyAgc = anyMs = false;
        msSplitMask = 0;
        for (auto& v : msStreams) v.clear();
        parMsStreams.clear();
        for (int q = 0; q < nSeq; ++q)
        {
            const int st = q / cfg.n_channels, ch = q % cfg.n_channels;
            const EqSet& e = eqSets[cfg.shared_eq ? 0 : (size_t) st];
            // An active Mid/Side band sends the reference to its node path (Processing.cpp:1037-1044), where BandNode::active
            // (with createBandNode's 0-dB skip) decides which bands run instead of EQCoeffCache::bandActive
            bool nodePath = false;
            for (int b = 0; b < CPQ_NUM_BANDS; ++b) nodePath |= e.active[b] && e.mode[b] >= 3;
            const uint8_t* on_ = (nodePath && e.hasNodeActive) ? e.nodeActive : e.active;
            unsigned m = 0;
            for (int b = 0; b < CPQ_NUM_BANDS; ++b)
            {
                if (!on_[b]) continue;
                const int mode = e.mode[b];
                // Processing.cpp:1239-1252: Stereo -> both; Left -> ch 0; Right -> ch 1
                const bool on = (mode == 0) || (mode == 1 && ch == 0) || (mode == 2 && ch == 1);
                if (on) m |= 1u << b;
                // Stereo mode on a stereo stream is processBandStereo (SSE); everything else goes through the scalar processBand
                if (on && (mode != 0 || cfg.n_channels < 2)) scalar[(size_t) q] |= 1u << b;
                if (mode >= 3 && ch == 0 && e.structure != 1) msStreams[b].push_back(st);   // ascending stream order
                if (mode >= 3 && ch == 0 && e.structure == 1 && (parMsStreams.empty() || parMsStreams.back() != st)) parMsStreams.push_back(st);
            }
            if (e.structure == 1) m |= 1u << 31;   // Parallel structure flag, read by eq_kernel<.., PAR>
            anyPar |= e.structure == 1;
            anyAgc |= e.agc != 0;
            mask[(size_t) q] = m;
            sos[(size_t) q] = cfg.shared_eq ? 0 : st;
        }
        CPQ_CUDA(eqc.ensure(host.size()));
        CPQ_CUDA(satDev.ensure(nSets));
        CPQ_CUDA(gainConst.ensure(nSets));
        CPQ_CUDA(bandMask.ensure((size_t) nSeq));
        CPQ_CUDA(scalarMask.ensure((size_t) nSeq));
        CPQ_CUDA(cudaMemcpyAsync(scalarMask.p, scalar.data(), scalar.size() * sizeof(unsigned), cudaMemcpyHostToDevice, stream));
        CPQ_CUDA(setOfSeq.ensure((size_t) nSeq));
        CPQ_CUDA(cudaMemcpyAsync(eqc.p, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        CPQ_CUDA(cudaMemcpyAsync(satDev.p, sat.data(), nSets * sizeof(double), cudaMemcpyHostToDevice, stream));
        CPQ_CUDA(cudaMemcpyAsync(gainConst.p, gc.data(), nSets * sizeof(double), cudaMemcpyHostToDevice, stream));
        CPQ_CUDA(cudaMemcpyAsync(bandMask.p, mask.data(), mask.size() * sizeof(unsigned), cudaMemcpyHostToDevice, stream));
        CPQ_CUDA(cudaMemcpyAsync(setOfSeq.p, sos.data(), sos.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        // Mid/Side bands: per band the compact list of streams, the scratch rows' band masks and parameter sets
        for (int b = 0; b < CPQ_NUM_BANDS; ++b)
        {
            const std::vector<int>& L = msStreams[b];
            if (L.empty()) continue;
            anyMs = true;
            msSplitMask |= 1u << b;
            std::vector<unsigned> mm(2 * L.size(), 0u);
            std::vector<int> ss(2 * L.size());
            for (size_t i = 0; i < L.size(); ++i)
            {
                const EqSet& e = eqSets[cfg.shared_eq ? 0 : (size_t) L[i]];
                mm[2 * i + (e.mode[b] == 3 ? 0 : 1)] = 1u << b;   // Mid -> row 0, Side -> row 1
                ss[2 * i] = ss[2 * i + 1] = cfg.shared_eq ? 0 : L[i];
            }
            CPQ_CUDA(msStreamsDev[b].ensure(L.size()));
            CPQ_CUDA(msMaskDev[b].ensure(mm.size()));
            CPQ_CUDA(msSetDev[b].ensure(ss.size()));
            CPQ_CUDA(cudaMemcpyAsync(msStreamsDev[b].p, L.data(), L.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
            CPQ_CUDA(cudaMemcpyAsync(msMaskDev[b].p, mm.data(), mm.size() * sizeof(unsigned), cudaMemcpyHostToDevice, stream));
            CPQ_CUDA(cudaMemcpyAsync(msSetDev[b].p, ss.data(), ss.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        }
        if (!parMsStreams.empty())
        {
            anyMs = true;
            const std::vector<int>& L = parMsStreams;
            std::vector<unsigned> mm(2 * L.size(), 0u);
            std::vector<int> ss(2 * L.size());
            for (size_t i = 0; i < L.size(); ++i)
            {
                const EqSet& e = eqSets[cfg.shared_eq ? 0 : (size_t) L[i]];
                const uint8_t* on_ = e.hasNodeActive ? e.nodeActive : e.active;   // a Mid/Side band is active: node path
                unsigned mMid = 1u << 31, mSide = 1u << 31;
                for (int b = 0; b < CPQ_NUM_BANDS; ++b)
                {
                    if (!on_[b]) continue;
                    if (e.mode[b] == 3) mMid |= 1u << b;
                    if (e.mode[b] == 4) mSide |= 1u << b;
                }
                mm[2 * i] = mMid;
                mm[2 * i + 1] = mSide;
                ss[2 * i] = ss[2 * i + 1] = cfg.shared_eq ? 0 : L[i];
            }
            CPQ_CUDA(parMsStreamsDev.ensure(L.size()));
            CPQ_CUDA(parMsMaskDev.ensure(mm.size()));
            CPQ_CUDA(parMsSetDev.ensure(ss.size()));
            CPQ_CUDA(cudaMemcpyAsync(parMsStreamsDev.p, L.data(), L.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
            CPQ_CUDA(cudaMemcpyAsync(parMsMaskDev.p, mm.data(), mm.size() * sizeof(unsigned), cudaMemcpyHostToDevice, stream));
            CPQ_CUDA(cudaMemcpyAsync(parMsSetDev.p, ss.data(), ss.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        }
        if (anyAgc)
        {
            std::vector<uint8_t> on((size_t) cfg.n_streams);
            for (int st = 0; st < cfg.n_streams; ++st) on[(size_t) st] = eqSets[cfg.shared_eq ? 0 : (size_t) st].agc ? 1 : 0;
            CPQ_CUDA(agcOnDev.ensure(on.size()));
            CPQ_CUDA(agcState.ensure(on.size() * 3));
            CPQ_CUDA(cudaMemcpyAsync(agcOnDev.p, on.data(), on.size(), cudaMemcpyHostToDevice, stream));
        }
        CPQ_CUDA(cudaStreamSynchronize(stream));
        // the compact stream lists may have changed -- unless the states were just imported for exactly these settings
        if (msState.p && !msImported) CPQ_CUDA(cudaMemsetAsync(msState.p, 0, msState.n * sizeof(double), stream));
        msImported = false;
        eqDirty = false;
        gainTabCallbacks = -1;
    }
    if (anyMs && cfg.n_channels != 2)
    {
        setError("Mid/Side bands need a stereo handle (n_channels = 2)");
        return CPQ_ERR_UNSUPPORTED;
    }
    // streaming continuation: a ramp that a previous call started goes on in this one (and a set whose ramp has settled at a
    // new gain keeps it: the table then holds that gain for every callback)
    bool anyEvents = false;
    for (auto& e : eqSets) anyEvents |= !e.events.empty() || (streaming && e.ramp.live);
    if (anyEvents && anyAgc)
    {
        setError("total-gain events and AGC in one handle: with AGC the reference never applies the total-gain ramp (Processing.cpp:1255-1274)");
        return CPQ_ERR_UNSUPPORTED;
    }
    haveGainTab = anyEvents;
    if (anyEvents && gainTabCallbacks != nCallbacks)
    {
        std::vector<double> tab(nSets * (size_t) nCallbacks * 2), one;
        const int steps = std::max(1, (int) (cfg.sample_rate * 0.05 + 0.5));   // SMOOTHING_TIME_SEC, computeTotalSteps
        for (size_t s = 0; s < nSets; ++s)
        {
            gainRampTable(eqSets[s].totalGain, steps, cfg.block_size, nCallbacks, eqSets[s].events, one, streaming ? &eqSets[s].ramp : nullptr);
            std::memcpy(tab.data() + s * (size_t) nCallbacks * 2, one.data(), one.size() * sizeof(double));
            if (streaming) eqSets[s].events.clear();   // consumed: at_callback counts from the call that follows the scheduling
        }
        CPQ_CUDA(gainTab.ensure(tab.size()));
        CPQ_CUDA(cudaMemcpyAsync(gainTab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        CPQ_CUDA(cudaStreamSynchronize(stream));
        gainTabCallbacks = streaming ? -1 : nCallbacks;   // a stream's table is this call's only
    }
    return CPQ_OK;
}

// ---- streaming continuation state ---------------------------------------------------------------------------------
// Geometry of the carried state, fixed by the layer plan:  input history 2 Pmax samples;  FDL rows Q - 1 per layer;  tail
// window = the frames Get may still read: delayLineReadAdd starts at max(readCursor, writeCursor - outputDelaySamples)
// (MKLNonUniformConvolver.cpp:1640-1688) and writeCursor trails the input by at most one distribution cycle.
cpq_status Engine::ensureStreamState()
{
    if (!planSet)
    {
        // EQ / output stages only: nothing of the convolver to carry
        histLen = 0;
        return CPQ_OK;
    }
    int pmax = 0;
    for (int li = 0; li < plan.numLayers; ++li) pmax = std::max(pmax, plan.layers[li].partSize);
    const int wantHist = 2 * pmax;
    const bool fresh = histLen != wantHist || !inHistBuf[0].p;
    histLen = wantHist;
    const int B = cfg.block_size;
    for (int li = 0; li < plan.numLayers; ++li)
    {
        const LayerPlan& l = plan.layers[li];
        fdlRows[li] = l.numPartsIR - 1;
        const int cbs = (l.numPartsIR + std::max(1, l.partsPerCallback) - 1) / std::max(1, l.partsPerCallback);
        // layer 0 carries output only when the host block differs from its partition: the reference's output ring then holds up
        // to one partition plus one callback of samples Get has not read yet
        carryFrames[li] = li == 0 ? (l.partSize != B ? 3 : 0)
                                  : (l.outputDelaySamples + l.partSize - 1) / l.partSize + (int) (((int64_t) cbs * B + l.partSize - 1) / l.partSize) + 2;
    }
    if (!fresh) return CPQ_OK;
    for (auto& b : inHistBuf) CPQ_CUDA(b.ensure((size_t) nSeq * histLen));
    for (int li = 0; li < plan.numLayers; ++li)
    {
        const LayerPlan& l = plan.layers[li];
        if (carryFrames[li] > 0) CPQ_CUDA(tailCarry[li].ensure((size_t) nSeq * carryFrames[li] * l.partSize));
    }
    return resetState();
}

// Room for `newFrames` more spectra behind the carried ones of layer li (see Engine::xs); `cont` = the carried rows are live
cpq_status Engine::prepareFdl(int li, int64_t newFrames, bool cont)
{
    const int F = fdlRows[li], P = plan.layers[li].partSize;
    const size_t rowB = (size_t) P * sizeof(double2);
    const int need = 2 * F + (int) newFrames;   // the copy-back of F rows to the start never overlaps its source
    if (!xs[li].p || xCap[li] < need)
    {
        const int cap = need + (int) newFrames;   // a second call of this size fits before the first copy-back
        DevBuf<double2> nb;
        CPQ_CUDA(nb.ensure((size_t) nSeq * cap * P));
        if (F > 0)
        {
            if (cont && xs[li].p)
                CPQ_CUDA(cudaMemcpy2DAsync(nb.p, (size_t) cap * rowB, xs[li].p + (size_t) (xHead[li] - F) * P, (size_t) xCap[li] * rowB, (size_t) F * rowB,
                                           (size_t) nSeq, cudaMemcpyDeviceToDevice, stream));
            else
                CPQ_CUDA(cudaMemset2DAsync(nb.p, (size_t) cap * rowB, 0, (size_t) F * rowB, (size_t) nSeq, stream));
            CPQ_CUDA(cudaStreamSynchronize(stream));   // the old buffer is released below
        }
        std::swap(xs[li].p, nb.p);
        std::swap(xs[li].n, nb.n);
        xCap[li] = cap;
        xHead[li] = F;
    }
    else if (!cont)
    {
        if (F > 0) CPQ_CUDA(cudaMemset2DAsync(xs[li].p, (size_t) xCap[li] * rowB, 0, (size_t) F * rowB, (size_t) nSeq, stream));
        xHead[li] = F;
    }
    else if (xHead[li] + newFrames > xCap[li])
    {
        if (F > 0)
            CPQ_CUDA(cudaMemcpy2DAsync(xs[li].p, (size_t) xCap[li] * rowB, xs[li].p + (size_t) (xHead[li] - F) * P, (size_t) xCap[li] * rowB, (size_t) F * rowB,
                                       (size_t) nSeq, cudaMemcpyDeviceToDevice, stream));
        xHead[li] = F;
    }
    return CPQ_OK;
}

// MKLNonUniformConvolver::Reset (.cpp:1693) + EQProcessor state clear + PsychoacousticDither::reset for every stream
cpq_status Engine::resetState()
{
    CPQ_CUDA(cudaSetDevice(cfg.device));
    for (auto& b : inHistBuf)
        if (b.p) CPQ_CUDA(cudaMemsetAsync(b.p, 0, b.n * sizeof(double), stream));
    for (int li = 0; li < CPQ_MAX_LAYERS; ++li)
    {
        if (xs[li].p && fdlRows[li] > 0)   // Reset: an empty FDL at the start of the buffer
            CPQ_CUDA(cudaMemset2DAsync(xs[li].p, (size_t) xCap[li] * plan.layers[li].partSize * sizeof(double2), 0,
                                       (size_t) fdlRows[li] * plan.layers[li].partSize * sizeof(double2), (size_t) nSeq, stream));
        xHead[li] = fdlRows[li];
        if (tailCarry[li].p) CPQ_CUDA(cudaMemsetAsync(tailCarry[li].p, 0, tailCarry[li].n * sizeof(double), stream));
    }
    CPQ_CUDA(cudaMemsetAsync(ditherZ.p, 0, (size_t) nSeq * 12 * sizeof(double), stream));
    if (rngState.p && !rngSeedState.empty())
        CPQ_CUDA(cudaMemcpyAsync(rngState.p, rngSeedState.data(), rngSeedState.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    CPQ_CUDA(cudaMemsetAsync(stateOut.p, 0, (size_t) nSeq * CPQ_NUM_BANDS * 2 * sizeof(double), stream));
    if (postState.p) CPQ_CUDA(cudaMemsetAsync(postState.p, 0, postState.n * sizeof(double), stream));
    if (msState.p) CPQ_CUDA(cudaMemsetAsync(msState.p, 0, msState.n * sizeof(double), stream));
    for (auto& d : dryHist)
        if (d.p) CPQ_CUDA(cudaMemsetAsync(d.p, 0, d.n * sizeof(double), stream));
    CPQ_CUDA(cudaStreamSynchronize(stream));
    absCallback = 0;
    contValid = false;
    outerPending = false;
    for (auto& e : eqSets)
    {
        e.events.clear();
        e.ramp = GainRampState {};
    }
    gainTabCallbacks = -1;
    return CPQ_OK;
}

// Serialised state: header, then the buffers in a fixed order.  Only meaningful for a handle with the same configuration,
// impulses and EQ settings (the header carries the geometry to check that).
struct StateHeader
{
    uint64_t magic;          // "CPQSTAT2"
    int32_t nSeq, nStreams, histLen, numLayers;
    int32_t fdlRows[CPQ_MAX_LAYERS], carryFrames[CPQ_MAX_LAYERS], partSize[CPQ_MAX_LAYERS];
    int32_t contValid, hasPost, hasAgc, hasLim;
    int32_t hasMs, dryLen;
    int64_t absCallback;
};
static constexpr uint64_t kStateMagic = 0x3254415453515043ull;   // "CPQSTAT2"

size_t Engine::stateBytes() const
{
    size_t n = sizeof(StateHeader);
    n += (size_t) nSeq * histLen * sizeof(double);
    if (histLen > 0)
        for (int li = 0; li < plan.numLayers; ++li)
        {
            n += (size_t) nSeq * fdlRows[li] * plan.layers[li].partSize * sizeof(double2);
            n += (size_t) nSeq * carryFrames[li] * plan.layers[li].partSize * sizeof(double);
        }
    n += (size_t) nSeq * CPQ_NUM_BANDS * 2 * sizeof(double);     // EQ band states
    n += (size_t) nSeq * kEqPostStages * 2 * sizeof(double);     // output-stage states
    n += (size_t) cfg.n_streams * 3 * sizeof(double);            // AGC envelopes + gain
    n += (size_t) cfg.n_streams * sizeof(double);                // limiter envelope
    n += (size_t) nSeq * 12 * sizeof(double);                    // dither error history
    n += (size_t) nSeq * sizeof(unsigned long long);             // dither generator state
    n += msStateCount() * sizeof(double);                        // Mid / Side band states
    n += (size_t) nSeq * dryHistLen * sizeof(double);            // dry path's delay ring
    return n;
}

cpq_status Engine::exportState(void* dst, size_t bytes)
{
    if (!streaming)
    {
        setError("export_state: the handle is not in streaming mode (cpq_set_streaming)");
        return CPQ_ERR_NOT_READY;
    }
    cpq_status st = ensureStreamState();
    if (st != CPQ_OK) return st;
    if (!dst || bytes < stateBytes())
    {
        setError("export_state: buffer smaller than cpq_state_size()");
        return CPQ_ERR_INVALID;
    }
    CPQ_CUDA(cudaSetDevice(cfg.device));
    CPQ_CUDA(cudaStreamSynchronize(stream));
    StateHeader h {};
    h.magic = kStateMagic;
    h.nSeq = nSeq; h.nStreams = cfg.n_streams; h.histLen = histLen; h.numLayers = histLen > 0 ? plan.numLayers : 0;
    for (int li = 0; li < h.numLayers; ++li) { h.fdlRows[li] = fdlRows[li]; h.carryFrames[li] = carryFrames[li]; h.partSize[li] = plan.layers[li].partSize; }
    h.contValid = contValid ? 1 : 0;
    h.hasPost = postState.p ? 1 : 0; h.hasAgc = agcState.p ? 1 : 0; h.hasLim = limEnv.p ? 1 : 0;
    h.hasMs = msState.p ? 1 : 0;
    h.dryLen = dryHistLen;
    h.absCallback = absCallback;
    char* o = static_cast<char*>(dst);
    std::memcpy(o, &h, sizeof(h));
    o += sizeof(h);
    auto put = [&](const void* dev, size_t n, bool present) -> cudaError_t {
        cudaError_t e = cudaSuccess;
        if (present && n) e = cudaMemcpy(o, dev, n, cudaMemcpyDeviceToHost);
        else std::memset(o, 0, n);
        o += n;
        return e;
    };
    CPQ_CUDA(put(inHistCur(), (size_t) nSeq * histLen * sizeof(double), histLen > 0));
    for (int li = 0; li < h.numLayers; ++li)
    {
        {
            // the FDL rows of every sequence, gathered out of the persistent buffers
            const size_t rowB = (size_t) plan.layers[li].partSize * sizeof(double2), n = (size_t) nSeq * fdlRows[li] * rowB;
            if (xs[li].p && n)
                CPQ_CUDA(cudaMemcpy2D(o, (size_t) fdlRows[li] * rowB, xs[li].p + (size_t) (xHead[li] - fdlRows[li]) * plan.layers[li].partSize,
                                      (size_t) xCap[li] * rowB, (size_t) fdlRows[li] * rowB, (size_t) nSeq, cudaMemcpyDeviceToHost));
            else
                std::memset(o, 0, n);
            o += n;
        }
        CPQ_CUDA(put(tailCarry[li].p, (size_t) nSeq * carryFrames[li] * plan.layers[li].partSize * sizeof(double), carryFrames[li] > 0));
    }
    CPQ_CUDA(put(stateOut.p, (size_t) nSeq * CPQ_NUM_BANDS * 2 * sizeof(double), true));
    CPQ_CUDA(put(postState.p, (size_t) nSeq * kEqPostStages * 2 * sizeof(double), postState.p != nullptr));
    CPQ_CUDA(put(agcState.p, (size_t) cfg.n_streams * 3 * sizeof(double), agcState.p != nullptr));
    CPQ_CUDA(put(limEnv.p, (size_t) cfg.n_streams * sizeof(double), limEnv.p != nullptr));
    CPQ_CUDA(put(ditherZ.p, (size_t) nSeq * 12 * sizeof(double), true));
    CPQ_CUDA(put(rngState.p, (size_t) nSeq * sizeof(unsigned long long), rngState.p != nullptr));
    CPQ_CUDA(put(msState.p, msStateCount() * sizeof(double), msState.p != nullptr));
    CPQ_CUDA(put(dryHist[dryHistSel].p, (size_t) nSeq * dryHistLen * sizeof(double), dryHistLen > 0));
    return CPQ_OK;
}

cpq_status Engine::importState(const void* src, size_t bytes)
{
    if (!streaming)
    {
        setError("import_state: the handle is not in streaming mode (cpq_set_streaming)");
        return CPQ_ERR_NOT_READY;
    }
    cpq_status st = ensureStreamState();
    if (st != CPQ_OK) return st;
    StateHeader h {};
    if (!src || bytes < sizeof(h))
    {
        setError("import_state: short buffer");
        return CPQ_ERR_INVALID;
    }
    std::memcpy(&h, src, sizeof(h));
    const size_t need = stateBytes() - (size_t) nSeq * dryHistLen * sizeof(double) + (size_t) nSeq * std::max(h.dryLen, 0) * sizeof(double);
    bool ok = h.magic == kStateMagic && h.nSeq == nSeq && h.nStreams == cfg.n_streams && h.histLen == histLen && bytes >= need &&
              h.numLayers == (histLen > 0 ? plan.numLayers : 0);
    for (int li = 0; ok && li < h.numLayers; ++li)
        ok = h.fdlRows[li] == fdlRows[li] && h.carryFrames[li] == carryFrames[li] && h.partSize[li] == plan.layers[li].partSize;
    if (!ok)
    {
        setError("import_state: the blob was exported from a handle with another configuration or layer plan");
        return CPQ_ERR_GEOMETRY;
    }
    CPQ_CUDA(cudaSetDevice(cfg.device));
    CPQ_CUDA(cudaStreamSynchronize(stream));
    const char* o = static_cast<const char*>(src) + sizeof(h);
    auto get = [&](void* dev, size_t n, bool present) -> cudaError_t {
        cudaError_t e = cudaSuccess;
        if (present && n) e = cudaMemcpy(dev, o, n, cudaMemcpyHostToDevice);
        o += n;
        return e;
    };
    CPQ_CUDA(get(inHistCur(), (size_t) nSeq * histLen * sizeof(double), histLen > 0));
    for (int li = 0; li < h.numLayers; ++li)
    {
        {
            const size_t rowB = (size_t) plan.layers[li].partSize * sizeof(double2), n = (size_t) nSeq * fdlRows[li] * rowB;
            cpq_status stf = prepareFdl(li, 1, false);   // (allocates on a fresh handle; the rows land at the start of the buffer)
            if (stf != CPQ_OK) return stf;
            CPQ_CUDA(cudaStreamSynchronize(stream));   // its clears run on the engine stream, the copy below does not
            if (n)
                CPQ_CUDA(cudaMemcpy2D(xs[li].p, (size_t) xCap[li] * rowB, o, (size_t) fdlRows[li] * rowB, (size_t) fdlRows[li] * rowB, (size_t) nSeq,
                                      cudaMemcpyHostToDevice));
            o += n;
        }
        CPQ_CUDA(get(tailCarry[li].p, (size_t) nSeq * carryFrames[li] * plan.layers[li].partSize * sizeof(double), carryFrames[li] > 0));
    }
    CPQ_CUDA(get(stateOut.p, (size_t) nSeq * CPQ_NUM_BANDS * 2 * sizeof(double), true));
    if (h.hasPost) CPQ_CUDA(postState.ensure((size_t) nSeq * kEqPostStages * 2));
    CPQ_CUDA(get(postState.p, (size_t) nSeq * kEqPostStages * 2 * sizeof(double), h.hasPost != 0));
    if (h.hasAgc) CPQ_CUDA(agcState.ensure((size_t) cfg.n_streams * 3));
    CPQ_CUDA(get(agcState.p, (size_t) cfg.n_streams * 3 * sizeof(double), h.hasAgc != 0));
    if (h.hasLim) CPQ_CUDA(limEnv.ensure((size_t) cfg.n_streams));
    CPQ_CUDA(get(limEnv.p, (size_t) cfg.n_streams * sizeof(double), h.hasLim != 0));
    CPQ_CUDA(get(ditherZ.p, (size_t) nSeq * 12 * sizeof(double), true));
    CPQ_CUDA(get(rngState.p, (size_t) nSeq * sizeof(unsigned long long), rngState.p != nullptr));
    if (h.hasMs) CPQ_CUDA(msState.ensure(msStateCount()));
    CPQ_CUDA(get(msState.p, msStateCount() * sizeof(double), h.hasMs != 0));
    if (h.dryLen > 0)
    {
        for (auto& d : dryHist) CPQ_CUDA(d.ensure((size_t) nSeq * h.dryLen));
        dryHistLen = h.dryLen;
        CPQ_CUDA(get(dryHist[dryHistSel].p, (size_t) nSeq * h.dryLen * sizeof(double), true));
    }
    msImported = h.hasMs != 0 && eqDirty;   // the first process call of a fresh handle uploads the band settings: keep the states
    absCallback = h.absCallback;
    contValid = h.contValid != 0;
    gplanCallbacks = -1;
    return CPQ_OK;
}

// Streaming calls need the plan of callbacks [cb0, cb0 + n) of a stream that may have been running for hours: the integer
// state machine is run ahead to a horizon and only again when the stream passes it (one pass per 65 536 callbacks, ~12 minutes of
// audio at block 512), not from callback 0 in every call.
cpq_status Engine::ensureGatherHost(int64_t nCallbacks)
{
    if (gplanHostCallbacks >= nCallbacks) return CPQ_OK;
    const int64_t horizon = (nCallbacks + 65536 + 65535) / 65536 * 65536;
    simulateCallbacks(plan, horizon, gplan);
    gplanHostCallbacks = horizon;
    gplanCallbacks = -1;   // the device copies (one-shot form) no longer match
    for (int li = 1; li < plan.numLayers; ++li) layer[li].hasBlockMap = !gplan.blockIdentity[li];
    return CPQ_OK;
}

cpq_status Engine::ensureGather(int64_t nCallbacks)
{
    if (gplanCallbacks == nCallbacks) return CPQ_OK;
    simulateCallbacks(plan, nCallbacks, gplan);
    gplanHostCallbacks = nCallbacks;
    if (!gplan.l0Identity)
    {
        CPQ_CUDA(layer[0].tailSrc.ensure((size_t) nCallbacks));
        CPQ_CUDA(layer[0].l0Count.ensure((size_t) nCallbacks));
        CPQ_CUDA(cudaMemcpyAsync(layer[0].tailSrc.p, gplan.l0Src.data(), (size_t) nCallbacks * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
        CPQ_CUDA(cudaMemcpyAsync(layer[0].l0Count.p, gplan.l0Count.data(), (size_t) nCallbacks * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    }
    for (int li = 1; li < plan.numLayers; ++li)
    {
        LayerDev& L = layer[li];
        CPQ_CUDA(L.tailSrc.ensure((size_t) nCallbacks));
        CPQ_CUDA(cudaMemcpyAsync(L.tailSrc.p, gplan.tailSrc[li].data(), (size_t) nCallbacks * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
        L.hasBlockMap = !gplan.blockIdentity[li];
        if (L.hasBlockMap)
        {
            CPQ_CUDA(L.blockMap.ensure(gplan.blockOfStream[li].size()));
            CPQ_CUDA(cudaMemcpyAsync(L.blockMap.p, gplan.blockOfStream[li].data(), gplan.blockOfStream[li].size() * sizeof(int32_t),
                                     cudaMemcpyHostToDevice, stream));
        }
    }
    CPQ_CUDA(cudaStreamSynchronize(stream));
    gplanCallbacks = nCallbacks;
    return CPQ_OK;
}

cpq_status Engine::launchEq(EqArgs& a)
{
    // one CTA per (sequence, kEqTile-sample tile); tiles of a sequence are chained through (sequence, tile, band) records
    // whose payload is its own flag: all-ones = not written yet
    a.nRuns = (int) ((a.T + kEqTile - 1) / kEqTile);
    // fewer sequences than SMs: tiles of the same sequence fill the machine and the links between them are the critical path
    static const int lbEnv = [] { const char* e = getenv("CPQ_EQ_LOOKBACK"); return e ? atoi(e) : -1; }();   // tuning knob
    const bool lookback = lbEnv >= 0 ? lbEnv != 0 : a.nSeq <= 32;
    a.chain.ticket = ticketFault.p;
    a.fault = ticketFault.p + 1;
    if ((a.doEq || a.postMask) && a.nRuns > 1)
    {
        const size_t n = (size_t) a.nSeq * a.nRuns * kEqStages * 2;
        CPQ_CUDA(chainRec.ensure(n));
        CPQ_CUDA(cudaMemsetAsync(chainRec.p, 0xff, n * sizeof(double), stream));
    }
    a.chain.rec = reinterpret_cast<double2*>(chainRec.p);
    CPQ_CUDA(cudaMemsetAsync(ticketFault.p, 0, sizeof(unsigned), stream));
    const unsigned grid = (unsigned) a.nSeq * (unsigned) a.nRuns;
    if (lookback)
    {
        if (a.doEq && anyPar) eq_kernel<true, true, true, true><<<grid, kEqThreads, eqSmemBytes(true, true), stream>>>(a);
        else if (a.sumsqIn || a.sumsqOut || a.nPeers > 0) eq_kernel<true, false, true, true><<<grid, kEqThreads, eqSmemBytes(true, true), stream>>>(a);
        else if (a.postMask || a.finalClamp || a.limFlag) eq_kernel<true, false, false, true><<<grid, kEqThreads, eqSmemBytes(true, true), stream>>>(a);
        else eq_kernel<false, false, false, true><<<grid, kEqThreads, eqSmemBytes(true, false), stream>>>(a);
    }
    else if (a.doEq && anyPar) eq_kernel<true, true, true><<<grid, kEqThreads, eqSmemBytes(false, true), stream>>>(a);
    else if (a.sumsqIn || a.sumsqOut || a.nPeers > 0) eq_kernel<true, false, true><<<grid, kEqThreads, eqSmemBytes(false, true), stream>>>(a);
    else if (a.postMask || a.finalClamp || a.limFlag) eq_kernel<true><<<grid, kEqThreads, eqSmemBytes(false, true), stream>>>(a);
    else eq_kernel<false><<<grid, kEqThreads, eqSmemBytes(false, false), stream>>>(a);
    ++launches;
    CPQ_CUDA(cudaGetLastError());
    return CPQ_OK;
}

// The EQ stage of one sequence chunk.  Serial / Parallel structure without AGC and without Mid/Side bands is the single
// fused launch.  Mid/Side bands split the band sequence: plain bands up to and including a split band run in one launch
// (a stream for which that band is Mid/Side has it masked out), then the band runs on the encoded Mid or Side row of the
// streams that want it (Processing.cpp:690-740).  AGC needs the statistics of whole callbacks before the gain can be
// applied (:1119-1131, :343-445): bands + statistics, the block-rate recurrence, then a launch that applies the gain
// ramp together with the output stages and the epilogue.
cpq_status Engine::runEq(EqArgs e, int s0, int ns)
{
    const bool agc = e.doEq && anyAgc, ms = e.doEq && anyMs;
    if (!agc && !ms) return launchEq(e);
    if (e.nPeers > 0)
    {
        // several launches follow: sum the peers' partials (and apply the deferred outer boundary) on their own first
        EqArgs w = e;
        w.doEq = 0;
        w.doGain = 0;
        w.gainTab = nullptr;
        w.postMask = 0;
        w.postStateOut = nullptr;
        w.doEpilogue = 0;
        w.finalClamp = 0;
        w.applyHeadroom = 0;
        cpq_status st = launchEq(w);
        if (st != CPQ_OK) return st;
        e.nPeers = 0;
        e.assemble = 0;
        e.nTail = 0;
        e.outer = 0;
    }
    const int nch = cfg.n_channels;
    const int st0 = s0 / nch, nst = ns / nch;
    double* sqIn = sumsq.p;
    double* sqOut = sumsq.p + (size_t) ns * e.nCallbacks;
    if (agc) CPQ_CUDA(cudaMemsetAsync(sumsq.p, 0, (size_t) 2 * ns * e.nCallbacks * sizeof(double), stream));
    auto strip = [](EqArgs& g) {   // bands only
        g.doGain = 0;
        g.gainTab = nullptr;
        g.postMask = 0;
        g.postStateOut = nullptr;
        g.doEpilogue = 0;
        g.finalClamp = 0;
        g.applyHeadroom = 0;
        g.limFlag = nullptr;
    };
    // streams of this chunk in the Parallel structure with Mid/Side bands (Processing.cpp:790-832): every band works on the
    // band input, so Mid and Side rows are encoded once from the input, all Mid bands run on the Mid row and all Side bands
    // on the Side row in one Parallel launch, the rows are reduced to the summed differences D_M, D_S, and after the plain
    // bands L += D_M + D_S, R += D_M - D_S
    const size_t p0 = (size_t) (std::lower_bound(parMsStreams.begin(), parMsStreams.end(), st0) - parMsStreams.begin());
    const size_t p1 = (size_t) (std::lower_bound(parMsStreams.begin(), parMsStreams.end(), st0 + nst) - parMsStreams.begin());
    const int parCnt = ms ? (int) (p1 - p0) : 0;
    MsArgs pm {};
    dim3 pg(1, 1);
    if (parCnt > 0)
    {
        if (e.assemble)
        {
            // the rows are encoded from the assembled convolver output: assemble first, on its own
            EqArgs w = e;
            strip(w);
            w.doEq = 0;
            cpq_status st = launchEq(w);
            if (st != CPQ_OK) return st;
            e.assemble = 0;
            e.nTail = 0;
            e.outer = 0;
        }
        pm.io = e.io;
        pm.ioStride = e.ioStride;
        pm.ms = parMsScratch.p;
        pm.streams = parMsStreamsDev.p + p0;
        pm.streamBase = st0;
        pm.T = e.T;
        pg = dim3((unsigned) std::min<int64_t>(64, ((e.T + 1) / 2 + 255) / 256), (unsigned) parCnt);
        ms_kernel<0><<<pg, 256, 0, stream>>>(pm);
        ++launches;
        EqArgs q = e;
        strip(q);
        q.io = parMsScratch.p;
        q.nSeq = 2 * parCnt;
        q.bandMask = parMsMaskDev.p + 2 * p0;
        q.scalarMask = nullptr;
        q.scalarAll = (1u << CPQ_NUM_BANDS) - 1u;   // Mid / Side rows: processBand
        q.setOfSeq = parMsSetDev.p + 2 * p0;
        q.stateOut = (streaming && msState.p) ? msRegion(CPQ_NUM_BANDS) + (size_t) 2 * p0 * CPQ_NUM_BANDS * 2 : nullptr;
        q.stateIn = e.stateIn ? q.stateOut : nullptr;
        cpq_status st = launchEq(q);
        if (st != CPQ_OK) return st;
        ms_kernel<2><<<pg, 256, 0, stream>>>(pm);
        ++launches;
        CPQ_CUDA(cudaGetLastError());
    }
    const bool needFinal = agc || parCnt > 0;   // the gain has to wait for whole-callback statistics / for the Mid-Side differences
    int lo = 0;
    for (int b = 0; b <= CPQ_NUM_BANDS; ++b)
    {
        const bool final_ = b == CPQ_NUM_BANDS;
        if (!final_ && !(ms && ((msSplitMask >> b) & 1u))) continue;
        // plain bands lo..b (b included: it is masked out per sequence where it is Mid/Side)
        EqArgs g = e;
        const int hi = final_ ? CPQ_NUM_BANDS - 1 : b;
        g.bandSelect = hi >= lo ? (((1u << (hi + 1)) - 1u) & ~((1u << lo) - 1u)) : 0u;
        g.bandSelectPar = lo == 0 ? (1u << CPQ_NUM_BANDS) - 1u : 0u;   // Parallel sequences run all their bands in the first launch
        if (lo > 0)
        {
            g.assemble = 0;
            g.nTail = 0;
            g.outer = 0;
        }
        if (!final_ || needFinal) strip(g);
        if (agc && lo == 0) g.sumsqIn = sqIn;
        if (agc && final_ && parCnt == 0) g.sumsqOut = sqOut;
        cpq_status st = launchEq(g);
        if (st != CPQ_OK) return st;
        lo = b + 1;
        if (final_) break;
        // the Mid/Side band b of this chunk's streams
        const std::vector<int>& L = msStreams[b];
        const size_t i0 = (size_t) (std::lower_bound(L.begin(), L.end(), st0) - L.begin());
        const size_t i1 = (size_t) (std::lower_bound(L.begin(), L.end(), st0 + nst) - L.begin());
        if (i1 <= i0) continue;
        const int cnt = (int) (i1 - i0);
        MsArgs m {};
        m.io = e.io;
        m.ioStride = e.ioStride;
        m.ms = msScratch.p;
        m.streams = msStreamsDev[b].p + i0;
        m.streamBase = st0;
        m.T = e.T;
        const dim3 mg((unsigned) std::min<int64_t>(64, ((e.T + 1) / 2 + 255) / 256), (unsigned) cnt);
        ms_kernel<0><<<mg, 256, 0, stream>>>(m);
        ++launches;
        EqArgs q = e;
        strip(q);
        q.assemble = 0;
        q.nTail = 0;
        q.outer = 0;
        q.io = msScratch.p;
        q.nSeq = 2 * cnt;
        q.bandMask = msMaskDev[b].p + 2 * i0;
        q.scalarMask = nullptr;
        q.scalarAll = (1u << CPQ_NUM_BANDS) - 1u;
        q.setOfSeq = msSetDev[b].p + 2 * i0;
        q.bandSelect = 1u << b;
        q.stateOut = (streaming && msState.p) ? msRegion(b) + (size_t) 2 * i0 * CPQ_NUM_BANDS * 2 : nullptr;
        q.stateIn = e.stateIn ? q.stateOut : nullptr;
        st = launchEq(q);
        if (st != CPQ_OK) return st;
        ms_kernel<1><<<mg, 256, 0, stream>>>(m);
        ++launches;
        CPQ_CUDA(cudaGetLastError());
    }
    if (parCnt > 0)
    {
        ms_kernel<3><<<pg, 256, 0, stream>>>(pm);
        ++launches;
        CPQ_CUDA(cudaGetLastError());
        if (agc)
        {
            // the output statistics need the Mid/Side differences: a statistics-only pass
            EqArgs w = e;
            strip(w);
            w.assemble = 0;
            w.nTail = 0;
            w.outer = 0;
            w.doEq = 0;
            w.sumsqOut = sqOut;
            cpq_status st = launchEq(w);
            if (st != CPQ_OK) return st;
        }
    }
    if (!needFinal) return CPQ_OK;
    if (!agc)
    {
        EqArgs f = e;   // total gain + output stages + epilogue
        f.assemble = 0;
        f.nTail = 0;
        f.outer = 0;
        f.doEq = 0;
        return launchEq(f);
    }
    AgcArgs a {};
    a.sumsqIn = sqIn;
    a.sumsqOut = sqOut;
    a.gainTab = agcTab.p;
    a.agcOn = agcOnDev.p + st0;
    a.gainConst = gainConst.p;
    a.setOfSeq = e.setOfSeq;
    a.stateOut = agcState.p + (size_t) st0 * 3;
    a.stateIn = e.stateIn ? a.stateOut : nullptr;   // streaming continuation
    a.nStreams = nst;
    a.nch = nch;
    a.nCallbacks = e.nCallbacks;
    const double n = (double) cfg.block_size, sr = cfg.sample_rate;
    a.blockN = n;
    a.attack = 1.0 - std::exp(-n / (sr * 0.2));    // AGC_ATTACK_TIME_SEC   (EQProcessor.h:167-169, Core.cpp:781-783)
    a.release = 1.0 - std::exp(-n / (sr * 2.0));   // AGC_RELEASE_TIME_SEC
    a.smooth = 1.0 - std::exp(-n / (sr * 0.2));    // AGC_SMOOTH_TIME_SEC
    agc_kernel<<<(unsigned) ((nst + 63) / 64), 64, 0, stream>>>(a);
    ++launches;
    CPQ_CUDA(cudaGetLastError());
    EqArgs f = e;   // gain ramp + output stages + epilogue
    f.assemble = 0;
    f.nTail = 0;
    f.outer = 0;
    f.doEq = 0;
    f.doGain = 1;
    f.gainTab = agcTab.p;
    f.gainBySeq = nch;
    return launchEq(f);
}

// Every error exit of the pipelined implementation leaves copies in flight on the three streams (H2D still reading, D2H
// still writing the caller's buffers): drain them before the status reaches the caller, who may free or reuse the buffers.
cpq_status Engine::processCore(double* dIo, int64_t stride, int64_t T, unsigned stages, double* const* hostPlanar, float* const* hostF)
{
    // The dither's shaper is one dependent chain per sequence (200 cycles per sample, 50 ms for 10 s of audio whatever the batch).
    // Run after a sequence chunk's EQ it trails the call by a whole chain; run per *time segment* through the streaming
    // continuation (segment boundaries on EQ tile boundaries, so every sample up to the quantiser is bit-identical to the one-shot
    // call) only the last segment's chain trails.  CPQ_DITHER_SEGMENTS=1 turns it off.
    // Measured at cfg4 (10 s, 2048 sequences): 1 / 2 / 3 / 4 segments 118 / 102 / 102 / 99 ms -- shorter segments leave the MAC's
    // 64-frame blocks of the 4096-sample layer half empty, which eats what the shorter trailing chain returns; segments of less
    // than 65536 samples are not worth it.
    static const int segEnv = [] { const char* e = getenv("CPQ_DITHER_SEGMENTS"); return e ? atoi(e) : 4; }();
    const int64_t segUnit = std::max<int64_t>(kEqTile, cfg.block_size);
    const int64_t nSeg = std::min<int64_t>(segEnv, T / 65536);
    const int64_t segLen = nSeg > 1 ? ((T + nSeg - 1) / nSeg + segUnit - 1) / segUnit * segUnit : T;
    bool noEvents = !haveGainTab;
    for (auto& e : eqSets) noEvents = noEvents && e.events.empty();
    // (segments are whole callbacks and whole EQ tiles: host blocks that do not divide 8192 samples stay in one piece)
    const bool segmented = segLen < T && segUnit % cfg.block_size == 0 && !streaming && !hostPlanar && !hostF && planSet && noEvents && ditherBits > 0 &&
                           (stages & CPQ_STAGE_CONV) && (stages & CPQ_STAGE_EPILOGUE) && T % cfg.block_size == 0;
    cpq_status st = CPQ_OK;
    bool done = false;
    if (segmented)
    {
        streaming = true;
        contValid = false;
        absCallback = 0;
        for (auto& e : eqSets) e.ramp = GainRampState {};
        cpq_timings sum {};
        done = true;
        for (int64_t off = 0; off < T && st == CPQ_OK; off += segLen)
        {
            const int64_t len = std::min(segLen, T - off);
            segUniTotal = T;
            segUniOffset = off;
            st = processCoreImpl(dIo + off, stride, len, stages, nullptr, nullptr, off + len < T);
            if ((st == CPQ_ERR_UNSUPPORTED || st == CPQ_ERR_OOM) && off == 0 && !workStarted)
            {
                cudaGetLastError();
                done = false;   // outside what the continuation carries (irregular plans, stream windows ...) or no room for its buffers: nothing was touched
                st = CPQ_OK;
                break;
            }
            sum.fft_fwd_ms += timings.fft_fwd_ms; sum.mac_ms += timings.mac_ms; sum.fft_inv_ms += timings.fft_inv_ms;
            sum.eq_ms += timings.eq_ms; sum.total_ms += timings.total_ms;
            sum.kernel_launches += timings.kernel_launches; sum.chunks += timings.chunks;
        }
        if (st != CPQ_OK)
            for (auto& d : sDither) cudaStreamSynchronize(d);
        streaming = false;
        contValid = false;
        absCallback = 0;
        forcedChunk = 0;
        segUniTotal = segUniOffset = 0;
        gplanCallbacks = -1;
        if (done) timings = sum;
    }
    if (!done) st = processCoreImpl(dIo, stride, T, stages, hostPlanar, hostF);
    if (st != CPQ_OK)
    {
        if (sIn) cudaStreamSynchronize(sIn);
        if (stream) cudaStreamSynchronize(stream);
        if (sOut) cudaStreamSynchronize(sOut);
        cudaGetLastError();
        if (streaming)
        {
            // a failed call may have advanced part of the carried state: the stream cannot be continued
            const std::string keep = err;
            resetState();
            err = keep + " (streaming state was reset)";
        }
    }
    return st;
}

cpq_status Engine::processCoreImpl(double* dIo, int64_t stride, int64_t T, unsigned stages, double* const* hostPlanar, float* const* hostF,
                                   bool deferSideStreams)
{
    const bool hostIO = hostPlanar || hostF;
    workStarted = false;
    // rows are 16-byte aligned (even stride); an odd T (odd host blocks: 441 x an odd number of callbacks) leaves one pad
    // sample at the end of each row, which the paired accesses of the small kernels may touch
    if (!dIo || T <= 0 || T > cfg.max_samples || T % cfg.block_size != 0 || stride < T + (T & 1) || (stride & 1))
    {
        setError("process: T must be a positive multiple of block_size <= max_samples; stride even and >= T");
        return CPQ_ERR_INVALID;
    }
    if ((stages & ~(CPQ_STAGE_FULL | CPQ_ORDER_EQ_THEN_CONV | CPQ_STAGE_INPUT)) || (stages & (CPQ_STAGE_FULL | CPQ_STAGE_INPUT)) == 0)
    {
        setError("process: bad stage mask");
        return CPQ_ERR_INVALID;
    }
    CPQ_CUDA(cudaSetDevice(cfg.device));
    const int B = cfg.block_size;
    const int64_t nCallbacks = T / B;
    const bool doConv = stages & CPQ_STAGE_CONV, doEq = stages & CPQ_STAGE_EQ, doEpi = stages & CPQ_STAGE_EPILOGUE;
    const int launches0 = (int) launches;
    timings = cpq_timings {};
    // streaming continuation: this call covers callbacks [cb0, cb0 + nCallbacks) of the stream since Reset
    const bool strm = streaming;
    const bool cont = strm && contValid;       // carried state exists
    const int64_t cb0 = strm ? absCallback : 0;

    if (doConv)
    {
        if (!planSet)
        {
            setError("process: no impulse set");
            return CPQ_ERR_NOT_READY;
        }
        for (int r = 0; r < nH; ++r)
            if (!haveImpulse[(size_t) r])
            {
                setError("process: cpq_set_impulse was not called for every stream-channel");
                return CPQ_ERR_NOT_READY;
            }
        cpq_status st = strm ? ensureGatherHost(cb0 + nCallbacks) : ensureGather(nCallbacks);
        if (st != CPQ_OK) return st;
    }
    if (doEq)
    {
        cpq_status st = uploadEq(nCallbacks);
        if (st != CPQ_OK) return st;
    }
    // linear output stages: OutputFilter between the EQ and the makeup gain, DC blocker inside the epilogue
    unsigned postMask = 0;
    if ((stages & CPQ_STAGE_OUTPUT_FILTER) && outCfg.filterEnabled) postMask |= 7u;
    if (doEpi && outCfg.dcCutoff > 0.0) postMask |= 8u;
    if (postMask)
    {
        cpq_status st = uploadPost();
        if (st != CPQ_OK) return st;
        postMask &= ~postIdentity;
        CPQ_CUDA(postState.ensure((size_t) nSeq * kEqPostStages * 2));
    }
    const bool doDither = doEpi && ditherBits > 0;
    const bool limiterOn = doEpi && limiterMs > 0.0;
    if (limiterOn)
    {
        CPQ_CUDA(limFlag.ensure((size_t) cfg.n_streams));
        CPQ_CUDA(limEnv.ensure((size_t) cfg.n_streams));
        // with dither the quantised signal is what the limiter sees: no detection in the EQ launch, every stream runs it
        CPQ_CUDA(cudaMemsetAsync(limFlag.p, doDither ? 0xff : 0, (size_t) cfg.n_streams * sizeof(unsigned), stream));
    }
    if (doDither && !cont) CPQ_CUDA(cudaMemsetAsync(ditherZ.p, 0, (size_t) this->nSeq * 12 * sizeof(double), stream));   // PsychoacousticDither::reset
    if (doDither && cont && segUniTotal == 0)
        for (auto& d : sDither) CPQ_CUDA(cudaStreamSynchronize(d));   // (paranoia: the carried history is read on the side streams)
    if (doDither && ditherRng && !cont)   // a fresh PsychoacousticDither per stream: generator state as constructed
        CPQ_CUDA(cudaMemcpyAsync(rngState.p, rngSeedState.data(), rngSeedState.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    if (doDither && !ditherRng && uniformsPerCh != (segUniTotal > 0 ? segUniTotal : T))
    {
        setError("process: dither enabled but cpq_set_dither_uniforms does not hold exactly T samples per channel");
        return CPQ_ERR_NOT_READY;
    }

    // partition range -> per-layer [qBegin, qEnd)
    int qb[CPQ_MAX_LAYERS] = {}, qe[CPQ_MAX_LAYERS] = {};
    bool fullRange = true;
    if (doConv)
    {
        int total = 0;
        for (int li = 0; li < plan.numLayers; ++li) total += plan.layers[li].numPartsIR;
        const int pb = std::max(0, partBegin), pe = partEnd < 0 ? total : std::min(partEnd, total);
        fullRange = (pb == 0 && pe == total);
        int base = 0;
        for (int li = 0; li < plan.numLayers; ++li)
        {
            const int Q = plan.layers[li].numPartsIR;
            qb[li] = clampv(0, Q, pb - base);
            qe[li] = clampv(0, Q, pe - base);
            base += Q;
        }
    }

    // ---- sequence chunking: bounded spectra workspace, and ~32 chunks so host copies overlap compute ----
    // stream window (partition-range sharding: a rank convolves every stream but finishes only its own)
    const int seqLo = std::min(std::max(winFirst, 0), cfg.n_streams) * cfg.n_channels;
    const int nSeqAll = this->nSeq;
    const int nSeq = (winCount < 0 ? nSeqAll - seqLo : std::min(winCount * cfg.n_channels, nSeqAll - seqLo));
    if (nSeq <= 0)
    {
        setError("process: empty stream window");
        return CPQ_ERR_INVALID;
    }
    // host block != L0 partition: the L0 output goes through the reference's output ring, i.e. to its own buffer, and the
    // EQ launch gathers it per callback; io keeps the input until then
    const bool l0Ring = doConv && !gplan.l0Identity;
    if (l0Ring && (directHead || !fullRange || nPeers > 0))
    {
        setError("process: a host block that is not a power of two cannot be combined with the direct-form head or partition-range sharding");
        return CPQ_ERR_UNSUPPORTED;
    }
    int64_t K[CPQ_MAX_LAYERS] = {};
    int64_t kOld[CPQ_MAX_LAYERS] = {};     // streaming: frames of the layer computed by earlier calls
    int64_t xRows[CPQ_MAX_LAYERS] = {};    // rows per sequence in the X workspace (carried FDL rows + new frames)
    if (strm)
    {
        if (!fullRange || nPeers > 0 || winFirst != 0 || winCount >= 0)
        {
            setError("process: streaming continuation does not cover partition-range sharding / stream windows");
            return CPQ_ERR_UNSUPPORTED;
        }
        if (doConv)
            for (int li = 1; li < plan.numLayers; ++li)
                if (!gplan.blockIdentity[li])
                {
                    setError("process: streaming continuation needs a regular layer plan (this plan drops tail blocks)");
                    return CPQ_ERR_UNSUPPORTED;
                }
        cpq_status st = ensureStreamState();
        if (st != CPQ_OK) return st;
        if (doEq && anyMs && !msState.p)
        {
            CPQ_CUDA(msState.ensure(msStateCount()));
            CPQ_CUDA(cudaMemsetAsync(msState.p, 0, msStateCount() * sizeof(double), stream));
        }
    }
    int chunk = nSeq;
    if (doConv)
    {
        size_t perSeq = 0;
        for (int li = 0; li < plan.numLayers; ++li)
        {
            const LayerPlan& l = plan.layers[li];
            if (strm)
            {
                // every frame whose input is complete at the end of this call, minus those earlier calls computed
                kOld[li] = cb0 * (int64_t) B / l.partSize;
                K[li] = (cb0 + nCallbacks) * (int64_t) B / l.partSize - kOld[li];
                xRows[li] = K[li];   // the chunk workspace only serves as scratch in streaming mode: the spectra live in Engine::xs
                perSeq += (size_t) (xRows[li] + K[li]) * l.partSize * sizeof(double2);
                if (li > 0 || l0Ring) perSeq += (size_t) (carryFrames[li] + K[li]) * l.partSize * sizeof(double);
            }
            else
            {
                K[li] = gplan.framesNeeded[li];
                xRows[li] = K[li];
                perSeq += (size_t) K[li] * l.partSize * sizeof(double2) * 2;
                if (li > 0 || l0Ring) perSeq += (size_t) K[li] * l.partSize * sizeof(double);
            }
        }
        chunk = (int) std::max<size_t>(1, std::min<size_t>((size_t) nSeq, cfg.workspace_bytes / std::max<size_t>(perSeq, 1)));
    }
    // host buffers: ~32 chunks so that the copies of one chunk run beside the kernels of another (short pipeline fill / drain) --
    // for calls large enough for that to matter (16 MB); a streaming call of a few callbacks is one chunk, one copy each way
    const size_t callBytes = (size_t) nSeq * (size_t) T * (hostF ? sizeof(float) : sizeof(double));
    if (hostIO && callBytes >= ((size_t) 16 << 20)) chunk = std::max(1, std::min(chunk, (nSeq + 31) / 32));
    if (segUniTotal > 0)
    {
        // segmented call: every segment uses the first one's chunks, so that a sequence's dither stays on one side stream
        if (forcedChunk > 0) chunk = std::min(chunk, forcedChunk);
        else forcedChunk = chunk;
    }
    // pageable caller rows go through pinned staging slots (HostStager): smaller chunks bound the pinned memory (6 slots)
    static const int stageThreadsEnv = [] { const char* e = getenv("CPQ_STAGE_THREADS"); return e ? atoi(e) : -1; }();   // 0 = let the driver stage
    // (small calls are left to the driver's own staging: starting the threads would cost more than they save)
    const bool pageable = hostPlanar && stageThreadsEnv != 0 && callBytes >= ((size_t) 32 << 20) &&
                          (hostRowIsPageable(hostPlanar[seqLo]) || hostRowIsPageable(hostPlanar[seqLo + nSeq - 1]));
    if (pageable) chunk = std::max(1, std::min(chunk, (nSeq + 63) / 64));
    if (limiterOn) chunk = std::max(cfg.n_channels, chunk / cfg.n_channels * cfg.n_channels);   // whole streams per chunk
    if (doEq && (anyAgc || anyMs))
    {
        // AGC statistics and Mid/Side bands couple the channels of a stream: whole streams per chunk, scratch rows bounded
        const int nch = cfg.n_channels;
        if (anyMs) chunk = (int) std::min<size_t>((size_t) chunk, std::max<size_t>((size_t) nch, cfg.workspace_bytes / ((size_t) stride * sizeof(double))));
        chunk = std::max(nch, chunk / nch * nch);
        if (anyMs) CPQ_CUDA(msScratch.ensure((size_t) chunk * stride));
        if (!parMsStreams.empty()) CPQ_CUDA(parMsScratch.ensure((size_t) chunk * stride));
        if (anyAgc)
        {
            CPQ_CUDA(sumsq.ensure((size_t) 2 * chunk * nCallbacks));
            CPQ_CUDA(agcTab.ensure((size_t) (chunk / nch) * nCallbacks * 2));
        }
    }
    if (doConv)
        for (int li = 0; li < plan.numLayers; ++li)
        {
            const LayerPlan& l = plan.layers[li];
            CPQ_CUDA(layer[li].X.ensure((size_t) chunk * xRows[li] * l.partSize));
            CPQ_CUDA(layer[li].Y.ensure((size_t) chunk * std::max<int64_t>(K[li], 1) * l.partSize));
            if (li > 0 || l0Ring) CPQ_CUDA(layer[li].tail.ensure((size_t) chunk * ((strm ? carryFrames[li] : 0) + K[li]) * l.partSize + 2));
        }
    if (strm && doConv)
    {
        // this call's slice of the gather plan, relative to the start of each layer's workspace stream
        // (= the carried samples, then the new frames): position p of the layer's output stream sits at p - base
        std::vector<int64_t> rel((size_t) nCallbacks);
        for (int li = l0Ring ? 0 : 1; li < plan.numLayers; ++li)
        {
            const int P = plan.layers[li].partSize;
            const int64_t base = (kOld[li] - carryFrames[li]) * (int64_t) P;
            const int64_t have = (int64_t) (carryFrames[li] + K[li]) * P;
            for (int64_t c = 0; c < nCallbacks; ++c)
            {
                // layer 0: the output ring's read position and count of the callback; tails: the delay line's read position
                const int64_t sp = li == 0 ? gplan.l0Src[(size_t) (cb0 + c)] : gplan.tailSrc[li][(size_t) (cb0 + c)];
                const int64_t cnt = li == 0 ? gplan.l0Count[(size_t) (cb0 + c)] : B;
                rel[(size_t) c] = sp < 0 ? -1 : sp - base;
                if (sp >= 0 && cnt > 0 && (sp - base < 0 || sp - base + cnt > have))
                {
                    setError("process: streaming continuation: the gather plan reads outside the carried tail window");
                    return CPQ_ERR_UNSUPPORTED;
                }
            }
            CPQ_CUDA(layer[li].tailSrc.ensure((size_t) nCallbacks));
            CPQ_CUDA(cudaMemcpyAsync(layer[li].tailSrc.p, rel.data(), (size_t) nCallbacks * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
            if (li == 0) CPQ_CUDA(layer[0].l0Count.ensure((size_t) nCallbacks));
            if (li == 0)
                CPQ_CUDA(cudaMemcpyAsync(layer[0].l0Count.p, gplan.l0Count.data() + cb0, (size_t) nCallbacks * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
            CPQ_CUDA(cudaStreamSynchronize(stream));   // rel is reused for the next layer
        }
        gplanCallbacks = -1;   // the device copies of the plan are call-relative: rebuild before any other use
    }
    const size_t nChunks = (size_t) ((nSeq + chunk - 1) / chunk);
    // event pool layout: [c*6 + 0..4] stage boundaries on the compute stream, [c*6 + 5] H2D done on the copy-in stream
    for (size_t i = 0; i < nChunks * 12 + 4; ++i) poolEvent(i);
    const size_t dBase = nChunks * 9 + 4;   // per chunk [0] EQ done (compute stream), [1] dither / limiter done (side stream)
    const size_t pBase = nChunks * 11 + 4;  // per chunk: D2H into the pinned staging slot done (pageable caller rows)
    std::unique_ptr<HostStager> stager;
    if (pageable)
    {
        const size_t slotElems = (size_t) chunk * T;
        CPQ_CUDA(stageIn.ensure(HostStager::kSlots * slotElems));
        CPQ_CUDA(stageOut.ensure(HostStager::kSlots * slotElems));
        stager = std::make_unique<HostStager>();
        stager->device = cfg.device;
        const int hw = (int) std::thread::hardware_concurrency();
        stager->nThreads = stageThreadsEnv > 0 ? stageThreadsEnv : std::max(2, std::min(16, hw / 2));   // measured on a 16-core box, cfg4: 4 threads 493 ms, 16 threads 390 ms, driver staging 1161 ms
        stager->T = T;
        stager->slotElems = slotElems;
        stager->rows = hostPlanar;
        stager->in = stageIn.p;
        stager->out = stageOut.p;
        for (size_t c = 0; c < nChunks; ++c)
        {
            const int s0 = seqLo + (int) c * chunk;
            stager->chunks.push_back({ s0, std::min(chunk, seqLo + nSeq - s0) });
            stager->evH2D.push_back(evPool[c * 6 + 5]);
            stager->evD2H.push_back(evPool[pBase + c]);
        }
        stager->start();
    }
    // float host buffers: per chunk [0] inbound conversion done, [1] outbound conversion done (compute stream), [2] D2H done
    const size_t fBase = nChunks * 6 + 4;
    const int64_t Tp = (T + 1) & ~(int64_t) 1;   // row pitch of the float staging slots (8-byte aligned rows)
    if (hostF)
    {
        CPQ_CUDA(f32In.ensure((size_t) 3 * chunk * Tp));
        CPQ_CUDA(f32Out.ensure((size_t) 2 * chunk * Tp));
    }
    cudaEvent_t evInBegin = evPool[nChunks * 6 + 0], evInEnd = evPool[nChunks * 6 + 1];
    cudaEvent_t evOutBegin = evPool[nChunks * 6 + 2], evOutEnd = evPool[nChunks * 6 + 3];

    cudaEventRecord(ev[0], stream);
    // Host rows at a constant pitch (one contiguous [n_seq][pitch] buffer) move as one 2-D copy per chunk;
    // otherwise row by row.  H2D runs at most two chunks ahead of compute so the driver's launch queue never
    // fills with copies (a full queue would stall the host before the first D2H is enqueued).
    ptrdiff_t hostPitch = 0;
    ptrdiff_t pitchF = 0;
    if (hostF && nSeqAll > 1)
    {
        pitchF = hostF[1] - hostF[0];
        for (int s = 2; s < nSeqAll && pitchF > 0; ++s)
            if (hostF[s] - hostF[s - 1] != pitchF) pitchF = 0;
        if (pitchF < T) pitchF = 0;
    }
    if (hostPlanar && nSeqAll > 1 && !pageable)
    {
        hostPitch = hostPlanar[1] - hostPlanar[0];
        for (int s = 2; s < nSeqAll && hostPitch > 0; ++s)
            if (hostPlanar[s] - hostPlanar[s - 1] != hostPitch) hostPitch = 0;
        if (hostPitch < T) hostPitch = 0;
    }
    auto enqueueH2D = [&](size_t c) -> cudaError_t {
        const int s0 = seqLo + (int) c * chunk, ns = std::min(chunk, seqLo + nSeq - s0);
        cudaError_t e = cudaSuccess;
        if (hostF)
        {
            // into staging slot c % 3, free once chunk c - 3 has been converted
            if (c >= 3) cudaStreamWaitEvent(sIn, evPool[fBase + (c - 3) * 3], 0);
            float* slot = f32In.p + (c % 3) * (size_t) chunk * Tp;
            if (pitchF > 0)
                e = cudaMemcpy2DAsync(slot, (size_t) Tp * sizeof(float), hostF[s0], (size_t) pitchF * sizeof(float), (size_t) T * sizeof(float),
                                      (size_t) ns, cudaMemcpyHostToDevice, sIn);
            else
                for (int s = s0; s < s0 + ns && e == cudaSuccess; ++s)
                    e = cudaMemcpyAsync(slot + (size_t) (s - s0) * Tp, hostF[s], (size_t) T * sizeof(float), cudaMemcpyHostToDevice, sIn);
        }
        else if (stager)
        {
            stager->waitStaged(c);   // the feeder threads have copied the chunk's rows into the pinned slot
            e = cudaMemcpy2DAsync(dIo + (size_t) s0 * stride, (size_t) stride * sizeof(double), stager->inSlot(c), (size_t) T * sizeof(double),
                                  (size_t) T * sizeof(double), (size_t) ns, cudaMemcpyHostToDevice, sIn);
        }
        else if (hostPitch == T && stride == T)   // both sides dense: one linear copy
            e = cudaMemcpyAsync(dIo + (size_t) s0 * stride, hostPlanar[s0], (size_t) ns * T * sizeof(double), cudaMemcpyHostToDevice, sIn);
        else if (hostPitch > 0)
            e = cudaMemcpy2DAsync(dIo + (size_t) s0 * stride, (size_t) stride * sizeof(double), hostPlanar[s0], (size_t) hostPitch * sizeof(double),
                                  (size_t) T * sizeof(double), (size_t) ns, cudaMemcpyHostToDevice, sIn);
        else
            for (int s = s0; s < s0 + ns && e == cudaSuccess; ++s)
                e = cudaMemcpyAsync(dIo + (size_t) s * stride, hostPlanar[s], (size_t) T * sizeof(double), cudaMemcpyHostToDevice, sIn);
        if (e == cudaSuccess) e = cudaEventRecord(evPool[c * 6 + 5], sIn);
        if (stager && e == cudaSuccess) stager->noteH2D(c);
        if (c + 1 == nChunks && e == cudaSuccess) e = cudaEventRecord(evInEnd, sIn);
        return e;
    };
    constexpr size_t kH2DAhead = 2;
    if (hostIO)
    {
        cudaStreamWaitEvent(sIn, ev[0], 0);
        cudaEventRecord(evInBegin, sIn);
        for (size_t c = 0; c < std::min(kH2DAhead, nChunks); ++c) CPQ_CUDA(enqueueH2D(c));
    }

    auto fillEqCommon = [&](EqArgs& a) {
        a.blockLog2 = (B & (B - 1)) == 0 ? ilog2(B) : -1;
        a.blockSize = B;
        a.T = T;
        a.ioStride = stride;
        a.doEq = doEq ? 1 : 0;
        a.eqc = eqc.p;
        a.sat = satDev.p;
        a.gainTab = (doEq && haveGainTab) ? gainTab.p : nullptr;
        a.doGain = a.doEq;
        a.bandSelect = (1u << CPQ_NUM_BANDS) - 1u;
        a.bandSelectPar = a.bandSelect;
        a.gainConst = gainConst.p;
        a.nCallbacks = nCallbacks;
        a.doEpilogue = doEpi ? 1 : 0;
        a.makeup = makeup;
        a.applyHeadroom = (doEpi && ditherBits <= 0) ? 1 : 0;
        a.postc = postc.p;
        a.postMask = postMask;
        // scrub (bit 0) + clamp (bit 1); with dither the dither kernel does both, with the limiter the limiter kernel clamps
        a.finalClamp = (doEpi && outCfg.finalClamp && ditherBits <= 0) ? (limiterOn ? 1 : 3) : 0;
        a.limFlag = nullptr;
        a.limDiv = cfg.n_channels;
        a.nPeers = 0;
        a.wetGain = equalPowerSin(cfg.conv_boundary == CPQ_CONV_OUTER ? mix : 1.0) * 1.0;   // CONVOLUTION_HEADROOM_GAIN = 1.0 (ConvolverProcessor.h:209)
    };
    const bool deferredOuter = !doConv && outerPending;
    // ConvolverProcessor::process with mix < 1 (Runtime.cpp:367-377): the dry path needs the convolver's input, kept in a
    // copy because the L0 inverse transform overwrites io; mix <= 0.001 is the dry-only fast path (no convolution at all)
    const bool outerMix = doConv && cfg.conv_boundary == CPQ_CONV_OUTER;
    const bool needsDry = outerMix && (convBypassed || mix < 0.999), dryOnly = outerMix && (convBypassed || !(mix > 0.001));
    if (needsDry && !fullRange)
    {
        setError("process: mix < 1 together with a partition range (the dry path belongs to the rank that owns the sum)");
        return CPQ_ERR_UNSUPPORTED;
    }
    // the direct-form head belongs to the rank that holds partition 0 of layer 0
    const bool direct = doConv && directHead && !dryOnly && qb[0] == 0 && qe[0] > 0;
    if (needsDry || direct) CPQ_CUDA(dryBuf.ensure((size_t) chunk * stride));
    const bool dryCarry = strm && needsDry && dryDelay > 0;
    if (dryCarry)
    {
        const bool fresh = dryHistLen != dryDelay || !dryHist[0].p;
        for (auto& d : dryHist) CPQ_CUDA(d.ensure((size_t) this->nSeq * dryDelay));
        dryHistLen = dryDelay;
        if (fresh || !cont) CPQ_CUDA(cudaMemsetAsync(dryHist[dryHistSel].p, 0, (size_t) this->nSeq * dryDelay * sizeof(double), stream));
    }
    // ProcessingOrder::EQThenConvolver (DSPCoreDouble.cpp:415-451): EQ (with its total-gain ramp) on the raw input, the
    // convolver input trim, then the convolver; the final launch then only assembles the layers and runs the output stages
    const bool eqFirst = (stages & CPQ_ORDER_EQ_THEN_CONV) && doConv && doEq;

    if (strm && doConv && !dryOnly && !cont)
    {
        // a stream that starts here starts from Reset, whatever an earlier stream left in the carried buffers
        if (inHistCur()) CPQ_CUDA(cudaMemsetAsync(inHistCur(), 0, (size_t) this->nSeq * histLen * sizeof(double), stream));
        for (int li = 0; li < plan.numLayers; ++li)
            if (tailCarry[li].p) CPQ_CUDA(cudaMemsetAsync(tailCarry[li].p, 0, tailCarry[li].n * sizeof(double), stream));
    }
    if (strm && doConv && !dryOnly)
        for (int li = 0; li < plan.numLayers; ++li)
        {
            cpq_status st = prepareFdl(li, K[li], cont);
            if (st != CPQ_OK) return st;
        }
    workStarted = true;   // from here on the caller's buffers are modified
    for (size_t c = 0; c < nChunks; ++c)
    {
        const int s0 = seqLo + (int) c * chunk, ns = std::min(chunk, seqLo + nSeq - s0);
        double* ioC = dIo + (size_t) s0 * stride;
        cudaEvent_t* ce = &evPool[c * 6];
        if (hostIO) cudaStreamWaitEvent(stream, evPool[c * 6 + 5], 0);
        const dim3 cvGrid((unsigned) std::min<int64_t>(128, ((T + 1) / 2 + 255) / 256), (unsigned) ns);
        if (hostF)
        {
            convert_kernel<true><<<cvGrid, 256, 0, stream>>>(f32In.p + (c % 3) * (size_t) chunk * Tp, ioC, stride, T);
            ++launches;
            cudaEventRecord(evPool[fBase + c * 3], stream);
        }
        cudaEventRecord(ce[0], stream);
        if (stages & CPQ_STAGE_INPUT)
        {
            InputArgs ia {};
            ia.io = ioC;
            ia.stride = stride;
            ia.T = T;
            ia.gain = inputGain;
            ia.applyGain = std::fabs(inputGain - 1.0) > 1e-9 ? 1 : 0;
            ia.block = B;
            ia.vecEnd = B / 4 * 4;
            input_kernel<<<dim3((unsigned) std::min<int64_t>(256, (T + 255) / 256), (unsigned) ns), 256, 0, stream>>>(ia);
            ++launches;
            CPQ_CUDA(cudaGetLastError());
        }
        if (eqFirst)
        {
            EqArgs p {};
            fillEqCommon(p);
            p.io = ioC;
            p.nSeq = ns;
            p.bandMask = bandMask.p + s0;
            p.scalarMask = scalarMask.p + s0;
            p.setOfSeq = setOfSeq.p + s0;
            p.stateOut = stateOut.p + (size_t) s0 * CPQ_NUM_BANDS * 2;
            if (cont) p.stateIn = p.stateOut;
            p.postMask = 0;
            p.finalClamp = 0;
            p.applyHeadroom = 0;
            p.doEpilogue = std::fabs(convInputTrim - 1.0) > 1e-12 ? 1 : 0;   // scaleBlockFallback(ptr, n, convolverInputTrimGain), :438-445
            p.makeup = convInputTrim;
            cpq_status st = runEq(p, s0, ns);
            if (st != CPQ_OK) return st;
        }
        if (needsDry || direct)
            CPQ_CUDA(cudaMemcpy2DAsync(dryBuf.p, (size_t) stride * sizeof(double), ioC, (size_t) stride * sizeof(double), (size_t) T * sizeof(double),
                                       (size_t) ns, cudaMemcpyDeviceToDevice, stream));
        if (doConv && !dryOnly)
        {
            // The three stages of a layer for sequences [q0, q0 + n) of this chunk.
            auto fwdLayer = [&](int li, int q0, int n) -> cpq_status {
                const LayerPlan& l = plan.layers[li];
                if (K[li] == 0 || qb[li] >= qe[li]) return CPQ_OK;
                FwdArgs a {};
                a.src = ioC + (size_t) q0 * stride;
                a.srcStride = stride;
                a.frameStart0 = -(int64_t) l.partSize;
                a.lo = 0;
                a.hi = T;
                a.halfOnly = 0;
                a.framesPerSeq = (int) K[li];
                a.totalFrames = (int64_t) n * K[li];
                a.out = layer[li].X.p + (size_t) q0 * xRows[li] * l.partSize;
                a.outFramesPerSeq = (int) K[li];
                a.outFrameOffset = 0;
                if (strm)
                {
                    // the frames of this call reach back up to 2 P samples before it: those come from the carried input history
                    // (histLen = 2 Pmax), read beside the call's input -- no concatenated copy
                    a.frameStart0 = (kOld[li] - 1) * (int64_t) l.partSize - cb0 * (int64_t) B;
                    a.lo = -(int64_t) histLen;
                    a.histEnd = inHistCur() + (size_t) (s0 + q0) * histLen + histLen;
                    a.histStride = histLen;
                    // the new spectra go behind the carried ones in the sequence's FDL buffer -- directly, except for the four-step
                    // transforms (P > 8192), whose scratch shares the geometry of their output: those write the chunk workspace
                    // and the rows are copied over
                    if (l.partSize <= 8192)
                    {
                        a.out = xs[li].p + (size_t) (s0 + q0) * xCap[li] * l.partSize;
                        a.outFramesPerSeq = xCap[li];
                        a.outFrameOffset = xHead[li];
                    }
                }
                a.tw = layer[li].tw.p;
                a.ptw = layer[li].ptw.p;
                a.scratch = layer[li].Y.p + (size_t) q0 * K[li] * l.partSize;   // free until the MAC writes it
                a.scale = 1.0;
                cpq_status stf = launchFwd(ilog2(l.partSize), a);
                if (stf == CPQ_OK && strm && l.partSize > 8192)
                {
                    const size_t rowB = (size_t) l.partSize * sizeof(double2);
                    CPQ_CUDA(cudaMemcpy2DAsync(xs[li].p + ((size_t) (s0 + q0) * xCap[li] + (size_t) xHead[li]) * l.partSize, (size_t) xCap[li] * rowB, a.out,
                                               (size_t) K[li] * rowB, (size_t) K[li] * rowB, (size_t) n, cudaMemcpyDeviceToDevice, stream));
                }
                return stf;
            };
            auto macLayer = [&](int li, int q0, int n) -> cpq_status {
                const LayerPlan& l = plan.layers[li];
                if (K[li] == 0) return CPQ_OK;
                if (qb[li] >= qe[li])
                {
                    // this rank holds no partition of the layer: its contribution is zero
                    if (li == 0) CPQ_CUDA(cudaMemset2DAsync(ioC + (size_t) q0 * stride, (size_t) stride * sizeof(double), 0, (size_t) T * sizeof(double), (size_t) n, stream));
                    else CPQ_CUDA(cudaMemsetAsync(layer[li].tail.p + (size_t) q0 * K[li] * l.partSize, 0, (size_t) n * K[li] * l.partSize * sizeof(double), stream));
                    return CPQ_OK;
                }
                MacArgs a {};
                a.X = layer[li].X.p + (size_t) q0 * xRows[li] * l.partSize;
                int64_t xPitch = xRows[li], xExtent = xRows[li];
                if (strm)
                {
                    // [carried | new] in place: a window of the sequence's FDL buffer that starts fdlRows before the new frames
                    a.X = xs[li].p + ((size_t) (s0 + q0) * xCap[li] + (size_t) (xHead[li] - fdlRows[li])) * l.partSize;
                    xPitch = xCap[li];
                    xExtent = xCap[li] - (xHead[li] - fdlRows[li]);
                }
                a.H = layer[li].H.p;
                a.Y = layer[li].Y.p + (size_t) q0 * K[li] * l.partSize;
                a.K = (int) K[li];
                a.P = l.partSize;
                a.Q = l.numPartsIR;
                a.qBegin = qb[li];
                a.qEnd = qe[li];
                a.hSeqStride = (int64_t) l.numPartsIR * l.partSize;
                a.hSeqMod = cfg.shared_ir ? cfg.n_channels : 0;
                a.seqBase = s0 + q0;
                if (!cfg.shared_ir) a.H += (size_t) (s0 + q0) * a.hSeqStride;   // H rows are absolute sequence indices
                // enough CTAs to fill the GPU a few times over, each amortising its H tile and ring warm-up over as many frames as possible
                const int binTiles = l.partSize / kMacBins;
                // blocks of 32 frames (four thread groups) when that leaves less of the last block empty than blocks of 64: short
                // calls in streaming mode, the dither's time segments
                const int nqL = qe[li] - qb[li];
                static const int g4Env = [] { const char* e = getenv("CPQ_MAC_G4"); return e ? atoi(e) : 1; }();   // tuning knob
                static const int tmaEnv = [] { const char* e = getenv("CPQ_MAC_TMA"); return e ? atoi(e) : 1; }();   // tuning knob
                const bool useTma = tmaEnv && nqL <= kMacTmaMaxTaps && macTmaSmemBytes(nqL) <= kMaxDynSmem;
                const bool g4 = useTma && g4Env && nqL - 1 <= 32 && (K[li] + 31) / 32 * 32 < (K[li] + 63) / 64 * 64;
                const int step = g4 ? 32 : kMacSuper;
                int fpc = (int) ((K[li] + step - 1) / step) * step;
                while (fpc > 2 * step && (int64_t) binTiles * n * ((K[li] + fpc - 1) / fpc) < 4 * 148 * 2) fpc = ((fpc / 2 + step - 1) / step) * step;
                static const int fpcEnv = [] { const char* e = getenv("CPQ_MAC_FPC"); return e ? atoi(e) : 0; }();   // tuning knob
                if (fpcEnv > 0 && fpcEnv % step == 0 && fpcEnv < fpc && K[li] <= 2 * step) fpc = fpcEnv;   // short layers only
                a.framesPerCta = fpc;
                a.hist = strm ? fdlRows[li] : 0;
                a.xRows = (int) xPitch;
                a.ringRows = fpc <= step ? (step + (a.qEnd - a.qBegin) - 1 + kMacKT - 1) / kMacKT * kMacKT : macRingRows(a.qEnd - a.qBegin);
                const size_t smem = macSmemBytes(a.qEnd - a.qBegin, a.ringRows);
                if (smem > kMaxDynSmem)
                {
                    setError("process: MAC tile exceeds the shared memory of one SM");
                    return CPQ_ERR_UNSUPPORTED;
                }
                dim3 grid((unsigned) binTiles, (unsigned) ((K[li] + fpc - 1) / fpc), (unsigned) n);
                // tensor-map staging (one TMA copy per 64-frame block, no CTA barrier in the frame loop) for filters of up to
                // 65 taps; longer ones (uniform-partition extension) keep the row-copy kernel
                MacTensorMap tmX, tmH;
                if (useTma &&
                    encodeSpectraMap(tmX, a.X, 2 * (uint64_t) l.partSize, (uint64_t) xExtent, (uint64_t) n, (uint32_t) step, (uint64_t) xPitch) &&
                    encodeSpectraMap(tmH, layer[li].H.p, 2 * (uint64_t) l.partSize, (uint64_t) l.numPartsIR, (uint64_t) nH, (uint32_t) (a.qEnd - a.qBegin)))
                {
                    if (g4) mac_tma_kernel<4><<<grid, kMacBins * 4, macTmaSmemBytes(a.qEnd - a.qBegin, 4), stream>>>(a, tmX, tmH);
                    else mac_tma_kernel<kMacGroups><<<grid, kMacThreads, macTmaSmemBytes(a.qEnd - a.qBegin), stream>>>(a, tmX, tmH);
                }
                else if (g4)
                {
                    setError("process: tensor-map encoding unavailable (driver entry point cuTensorMapEncodeTiled)");
                    return CPQ_ERR_CUDA;
                }
                else
                    mac_kernel<<<grid, kMacThreads, smem, stream>>>(a);
                ++launches;
                CPQ_CUDA(cudaGetLastError());
                return CPQ_OK;
            };
            auto invLayer = [&](int li, int q0, int n) -> cpq_status {
                const LayerPlan& l = plan.layers[li];
                const size_t tailPitch = (size_t) ((strm ? carryFrames[li] : 0) + K[li]) * l.partSize;
                if (strm && carryFrames[li] > 0 && (li > 0 || l0Ring))   // the delay line / output ring: samples computed by earlier calls that this call's callbacks still read
                    CPQ_CUDA(cudaMemcpy2DAsync(layer[li].tail.p + (size_t) q0 * tailPitch, tailPitch * sizeof(double),
                                               tailCarry[li].p + (size_t) (s0 + q0) * carryFrames[li] * l.partSize, (size_t) carryFrames[li] * l.partSize * sizeof(double),
                                               (size_t) carryFrames[li] * l.partSize * sizeof(double), (size_t) n, cudaMemcpyDeviceToDevice, stream));
                if (K[li] == 0 || qb[li] >= qe[li]) return CPQ_OK;
                InvArgs a {};
                a.in = layer[li].Y.p + (size_t) q0 * K[li] * l.partSize;
                a.framesPerSeq = (int) K[li];
                a.framesOut = (int) K[li];
                a.totalFrames = (int64_t) n * K[li];
                a.out = (li == 0 && !l0Ring) ? ioC + (size_t) q0 * stride : layer[li].tail.p + (size_t) q0 * tailPitch;
                a.outStride = (li == 0 && !l0Ring) ? stride : (int64_t) tailPitch;
                if (strm && (li > 0 || l0Ring)) a.out += (size_t) carryFrames[li] * l.partSize;
                a.tw = layer[li].tw.p;
                a.ptw = layer[li].ptw.p;
                a.scratch = layer[li].X.p + (size_t) q0 * xRows[li] * l.partSize;   // the MAC has consumed it
                return launchInv(ilog2(l.partSize), a);
            };
            // L0 in slices small enough for its spectra to stay in the L2 between the three launches (tuning knob; 0 = off)
            static const int l0Sub = [] { const char* e = getenv("CPQ_L0_SUB"); return e ? atoi(e) : 0; }();
            const bool sliceL0 = l0Sub > 0 && !strm && !l0Ring && plan.numLayers > 1 && l0Sub < ns;
            // ---- forward FFTs of every layer (all read the untouched input) ----
            for (int li = sliceL0 ? 1 : 0; li < plan.numLayers; ++li)
            {
                cpq_status st = fwdLayer(li, 0, ns);
                if (st != CPQ_OK) return st;
            }
            if (strm && histLen > 0)
            {
                // input history for the next call = the last histLen samples of [history | input], into the other buffer; now,
                // while io still holds the input (layer 0's inverse transform writes in place)
                double* nh = inHistBuf[inHistSel ^ 1].p + (size_t) s0 * histLen;
                const double* oh = inHistCur() + (size_t) s0 * histLen;
                const size_t hb = (size_t) histLen * sizeof(double);
                if (T >= histLen)
                    CPQ_CUDA(cudaMemcpy2DAsync(nh, hb, ioC + (T - histLen), (size_t) stride * sizeof(double), hb, (size_t) ns, cudaMemcpyDeviceToDevice, stream));
                else
                {
                    CPQ_CUDA(cudaMemcpy2DAsync(nh, hb, oh + T, hb, (size_t) (histLen - T) * sizeof(double), (size_t) ns, cudaMemcpyDeviceToDevice, stream));
                    CPQ_CUDA(cudaMemcpy2DAsync(nh + (histLen - T), hb, ioC, (size_t) stride * sizeof(double), (size_t) T * sizeof(double), (size_t) ns,
                                               cudaMemcpyDeviceToDevice, stream));
                }
            }
            cudaEventRecord(ce[1], stream);
            // ---- MAC ----
            if (sliceL0)
                for (int q0 = 0; q0 < ns; q0 += l0Sub)
                {
                    const int n = std::min(l0Sub, ns - q0);
                    cpq_status st = fwdLayer(0, q0, n);
                    if (st == CPQ_OK) st = macLayer(0, q0, n);
                    if (st == CPQ_OK) st = invLayer(0, q0, n);
                    if (st != CPQ_OK) return st;
                }
            for (int li = sliceL0 ? 1 : 0; li < plan.numLayers; ++li)
            {
                cpq_status st = macLayer(li, 0, ns);
                if (st != CPQ_OK) return st;
            }
            cudaEventRecord(ce[2], stream);
            // ---- inverse FFTs: L0 in place into io, tails into their stream buffers ----
            for (int li = sliceL0 ? 1 : 0; li < plan.numLayers; ++li)
            {
                cpq_status st = invLayer(li, 0, ns);
                if (st != CPQ_OK) return st;
            }
        }
        else
        {
            cudaEventRecord(ce[1], stream);
            cudaEventRecord(ce[2], stream);
        }
        if (direct)
        {
            DirectArgs d {};
            d.io = ioC;
            d.x = dryBuf.p;
            if (strm && cont && histLen >= 31)
            {
                // streaming continuation: the 31 samples before the call are the end of the input history this call started from
                d.histEnd = inHistCur() + (size_t) s0 * histLen + histLen;
                d.histStride = histLen;
            }
            d.taps = directTaps.p;
            d.stride = stride;
            d.T = T;
            d.hSeqMod = cfg.shared_ir ? cfg.n_channels : 0;
            d.seqBase = s0;
            direct_head_kernel<<<dim3((unsigned) std::min<int64_t>(128, (T + 255) / 256), (unsigned) ns), 256, 0, stream>>>(d);
            ++launches;
            CPQ_CUDA(cudaGetLastError());
        }
        cudaEventRecord(ce[3], stream);
        // ---- assembly + EQ + epilogue for this chunk ----
        EqArgs e {};
        fillEqCommon(e);
        e.io = ioC;
        e.nSeq = ns;
        if (!doConv && nPeers > 0)
        {
            // the ranks' convolver partials are summed (rank order) while this launch loads its tiles
            e.nPeers = nPeers;
            for (int p = 0; p < nPeers; ++p) e.peer[p] = peers[p] + (size_t) s0 * stride;
        }
        if (doConv && !dryOnly)
        {
            e.assemble = 1;
            e.nTail = plan.numLayers - 1;
            for (int li = 1; li < plan.numLayers; ++li)
            {
                e.tail[li - 1] = layer[li].tail.p;
                e.tailStride[li - 1] = (int64_t) ((strm ? carryFrames[li] : 0) + K[li]) * plan.layers[li].partSize;
                e.tailSrc[li - 1] = layer[li].tailSrc.p;
                e.blockMap[li - 1] = layer[li].hasBlockMap ? layer[li].blockMap.p : nullptr;
                e.tailPartLog2[li - 1] = ilog2(plan.layers[li].partSize);
                e.tailGain[li - 1] = plan.layers[li].gain;
            }
            e.outer = (cfg.conv_boundary == CPQ_CONV_OUTER && fullRange) ? 1 : 0;
            if (l0Ring)
            {
                e.l0 = layer[0].tail.p;
                e.l0Stride = (int64_t) ((strm ? carryFrames[0] : 0) + K[0]) * plan.layers[0].partSize;
                e.l0Src = layer[0].tailSrc.p;
                e.l0Count = layer[0].l0Count.p;
            }
        }
        else
        {
            e.assemble = deferredOuter ? 1 : 0;   // no tails; only the deferred outer-boundary scrub + wet gain
            e.nTail = 0;
            e.outer = deferredOuter ? 1 : 0;
        }
        if (eqFirst)
        {
            e.doEq = 0;
            e.doGain = 0;
            e.gainTab = nullptr;
        }
        e.bandMask = bandMask.p ? bandMask.p + s0 : nullptr;
        e.scalarMask = scalarMask.p ? scalarMask.p + s0 : nullptr;
        e.setOfSeq = setOfSeq.p ? setOfSeq.p + s0 : nullptr;   // absolute set indices
        e.stateOut = stateOut.p + (size_t) s0 * CPQ_NUM_BANDS * 2;
        e.postStateOut = postMask ? postState.p + (size_t) s0 * kEqPostStages * 2 : nullptr;
        if (cont)
        {
            e.stateIn = e.stateOut;
            e.postStateIn = e.postStateOut;
        }
        if (limiterOn && !doDither) e.limFlag = limFlag.p + s0 / cfg.n_channels;
        if (needsDry)
        {
            // assembly + scrub + wet gain on their own, then the dry path, then the remaining stages without assembly
            if (!dryOnly)
            {
                EqArgs w = e;
                w.doEq = 0;
                w.doGain = 0;
                w.gainTab = nullptr;
                w.postMask = 0;
                w.postStateOut = nullptr;
                w.doEpilogue = 0;
                w.finalClamp = 0;
                w.applyHeadroom = 0;
                cpq_status stw = launchEq(w);
                if (stw != CPQ_OK) return stw;
            }
            MixArgs m {};
            m.io = ioC;
            m.dry = dryBuf.p;
            m.stride = stride;
            m.T = T;
            m.delay = dryDelay;
            m.dryGain = equalPowerSin(1.0 - mix);
            m.dryOnly = dryOnly ? 1 : 0;
            if (dryCarry)
            {
                m.hist = dryHist[dryHistSel].p + (size_t) s0 * dryDelay;
                m.histOut = dryHist[dryHistSel ^ 1].p + (size_t) s0 * dryDelay;
            }
            mix_kernel<<<dim3((unsigned) std::min<int64_t>(256, (T + 255) / 256), (unsigned) ns), 256, 0, stream>>>(m);
            ++launches;
            CPQ_CUDA(cudaGetLastError());
            e.assemble = 0;
            e.nTail = 0;
            e.outer = 0;
        }
        if (doConv || doEq || doEpi || postMask || deferredOuter || e.nPeers > 0)   // an input-stage-only call has nothing left to do
        {
            cpq_status st = runEq(e, s0, ns);
            if (st != CPQ_OK) return st;
        }
        if (strm && doConv && !dryOnly)
            for (int li = l0Ring ? 0 : 1; li < plan.numLayers; ++li)
            {
                const LayerPlan& l = plan.layers[li];
                if (K[li] == 0) continue;   // nothing new: the carried window is unchanged
                CPQ_CUDA(cudaMemcpy2DAsync(tailCarry[li].p + (size_t) s0 * carryFrames[li] * l.partSize, (size_t) carryFrames[li] * l.partSize * sizeof(double),
                                           layer[li].tail.p + (size_t) K[li] * l.partSize, (size_t) (carryFrames[li] + K[li]) * l.partSize * sizeof(double),
                                           (size_t) carryFrames[li] * l.partSize * sizeof(double), (size_t) ns, cudaMemcpyDeviceToDevice, stream));
            }
        // The dither recurrence is one dependent chain per sequence (about 130 cycles per sample whatever the batch): on the
        // compute stream a chunk's dither would hold up the next chunk's transforms for T x 130 cycles with a handful of warps
        // busy, so it (and the limiter that follows it) runs on a side stream and the next chunks compute beside it.
        cudaStream_t post = stream;
        if (doDither)
        {
            cudaEventRecord(evPool[dBase + c * 2], stream);
            post = sDither[c % kDitherStreams];
            cudaStreamWaitEvent(post, evPool[dBase + c * 2], 0);
        }
        if (doDither)
        {
            DitherArgs d {};
            d.io = ioC;
            d.ioStride = stride;
            d.T = T;
            d.nSeq = ns;
            d.uniRow = segUniTotal > 0 ? segUniTotal : T;
            d.uniforms = ditherRng ? nullptr : (uniformsBorrowed ? uniformsBorrowed : uniforms.p) + ((size_t) s0 * d.uniRow + (size_t) segUniOffset) * 2;
            d.rng = ditherRng ? rngState.p + s0 : nullptr;
            ditherCoeffs(cfg.sample_rate, ditherBits, d.coeff);
            d.scale = 1.0 / std::pow(2.0, ditherBits - 1);
            d.invScale = std::pow(2.0, ditherBits - 1);
            d.z = ditherZ.p + (size_t) s0 * 12;
            d.finalClamp = outCfg.finalClamp ? (limiterOn ? 1 : 3) : 0;
            d.nch = cfg.n_channels;
            d.seqBase = s0;
            dither_kernel<<<(unsigned) ((ns + 31) / 32), kDitherThreads, kDitherSmemBytes, post>>>(d);
            ++launches;
            CPQ_CUDA(cudaGetLastError());
        }
        if (limiterOn)
        {
            LimiterArgs la {};
            la.io = ioC;
            la.stride = stride;
            la.T = T;
            la.nStreams = ns / cfg.n_channels;
            la.nch = cfg.n_channels;
            la.flag = limFlag.p + s0 / cfg.n_channels;
            la.release = std::exp(-1.0 / (cfg.sample_rate * limiterMs * 0.001));   // SimplePeakLimiter::prepare
            la.clamp = outCfg.finalClamp ? 1 : 0;
            la.envOut = limEnv.p + s0 / cfg.n_channels;
            la.envIn = cont ? la.envOut : nullptr;
            limiter_kernel<<<(unsigned) ((la.nStreams + 31) / 32), 32, 0, post>>>(la);
            ++launches;
            CPQ_CUDA(cudaGetLastError());
        }
        cudaEventRecord(ce[4], stream);
        if (doDither)
        {
            cudaEventRecord(evPool[dBase + c * 2 + 1], post);
            if (hostF) cudaStreamWaitEvent(stream, evPool[dBase + c * 2 + 1], 0);   // the outbound conversion reads the dithered signal
        }
        if (hostF)
        {
            // out through staging slot c % 2, free once chunk c - 2 has left the device
            if (c >= 2) cudaStreamWaitEvent(stream, evPool[fBase + (c - 2) * 3 + 2], 0);
            float* slot = f32Out.p + (c % 2) * (size_t) chunk * Tp;
            convert_kernel<false><<<cvGrid, 256, 0, stream>>>(slot, ioC, stride, T);
            ++launches;
            CPQ_CUDA(cudaGetLastError());
            cudaEventRecord(evPool[fBase + c * 3 + 1], stream);
            cudaStreamWaitEvent(sOut, evPool[fBase + c * 3 + 1], 0);
            if (c == 0) cudaEventRecord(evOutBegin, sOut);
            if (pitchF > 0)
                CPQ_CUDA(cudaMemcpy2DAsync(hostF[s0], (size_t) pitchF * sizeof(float), slot, (size_t) Tp * sizeof(float), (size_t) T * sizeof(float),
                                           (size_t) ns, cudaMemcpyDeviceToHost, sOut));
            else
                for (int s = s0; s < s0 + ns; ++s)
                    CPQ_CUDA(cudaMemcpyAsync(hostF[s], slot + (size_t) (s - s0) * Tp, (size_t) T * sizeof(float), cudaMemcpyDeviceToHost, sOut));
            cudaEventRecord(evPool[fBase + c * 3 + 2], sOut);
            if (c + kH2DAhead < nChunks) CPQ_CUDA(enqueueH2D(c + kH2DAhead));
        }
        else if (hostPlanar)
        {
            cudaStreamWaitEvent(sOut, doDither ? evPool[dBase + c * 2 + 1] : ce[4], 0);
            if (c == 0) cudaEventRecord(evOutBegin, sOut);
            if (stager)
            {
                if (c >= (size_t) HostStager::kSlots) stager->waitDrained(c - HostStager::kSlots);   // the slot's previous content has reached the caller
                CPQ_CUDA(cudaMemcpy2DAsync(stager->outSlot(c), (size_t) T * sizeof(double), dIo + (size_t) s0 * stride, (size_t) stride * sizeof(double),
                                           (size_t) T * sizeof(double), (size_t) ns, cudaMemcpyDeviceToHost, sOut));
                CPQ_CUDA(cudaEventRecord(evPool[pBase + c], sOut));
                stager->noteD2H(c);
            }
            else if (hostPitch == T && stride == T)
                CPQ_CUDA(cudaMemcpyAsync(hostPlanar[s0], dIo + (size_t) s0 * stride, (size_t) ns * T * sizeof(double), cudaMemcpyDeviceToHost, sOut));
            else if (hostPitch > 0)
                CPQ_CUDA(cudaMemcpy2DAsync(hostPlanar[s0], (size_t) hostPitch * sizeof(double), dIo + (size_t) s0 * stride, (size_t) stride * sizeof(double),
                                           (size_t) T * sizeof(double), (size_t) ns, cudaMemcpyDeviceToHost, sOut));
            else
                for (int s = s0; s < s0 + ns; ++s)
                    CPQ_CUDA(cudaMemcpyAsync(hostPlanar[s], dIo + (size_t) s * stride, (size_t) T * sizeof(double), cudaMemcpyDeviceToHost, sOut));
            if (c + kH2DAhead < nChunks) CPQ_CUDA(enqueueH2D(c + kH2DAhead));
        }
    }
    if (doConv) outerPending = (cfg.conv_boundary == CPQ_CONV_OUTER && !fullRange);
    else outerPending = false;
    if (strm)
    {
        absCallback += nCallbacks;
        contValid = true;
        if (dryCarry) dryHistSel ^= 1;
        if (doConv && !dryOnly && histLen > 0) inHistSel ^= 1;
        if (doConv && !dryOnly)
            for (int li = 0; li < plan.numLayers; ++li) xHead[li] += (int) K[li];   // the FDL now ends with this call's last frames
    }

    if (doDither && !deferSideStreams)
        for (size_t c = 0; c < nChunks; ++c) cudaStreamWaitEvent(stream, evPool[dBase + c * 2 + 1], 0);
    if (hostIO)
    {
        cudaEventRecord(evOutEnd, sOut);
        cudaStreamWaitEvent(stream, evOutEnd, 0);
    }
    cudaEventRecord(ev[6], stream);
    CPQ_CUDA(cudaEventSynchronize(ev[6]));
    if (stager)
    {
        stager->waitDrained(nChunks - 1);   // the drain threads take the chunks in order: the last one out means all are out
        stager.reset();
    }
    unsigned tf[2] = {};
    CPQ_CUDA(cudaMemcpy(tf, ticketFault.p, sizeof(tf), cudaMemcpyDeviceToHost));
    cudaEventElapsedTime(&timings.total_ms, ev[0], ev[6]);
    for (size_t c = 0; c < nChunks; ++c)
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evPool[c * 6 + 0], evPool[c * 6 + 1]); timings.fft_fwd_ms += ms;
        cudaEventElapsedTime(&ms, evPool[c * 6 + 1], evPool[c * 6 + 2]); timings.mac_ms += ms;
        cudaEventElapsedTime(&ms, evPool[c * 6 + 2], evPool[c * 6 + 3]); timings.fft_inv_ms += ms;
        cudaEventElapsedTime(&ms, evPool[c * 6 + 3], evPool[c * 6 + 4]); timings.eq_ms += ms;
    }
    if (hostIO)
    {
        cudaEventElapsedTime(&timings.h2d_ms, evInBegin, evInEnd);
        cudaEventElapsedTime(&timings.d2h_ms, evOutBegin, evOutEnd);
    }
    timings.chunks = (int) nChunks;
    timings.kernel_launches = (int) launches - launches0;
    if (std::getenv("CPQ_TRACE"))
    {
        // per-chunk timeline relative to the start of the call: H2D done, compute begin/end
        for (size_t c = 0; c < nChunks; ++c)
        {
            float tin = 0.f, tb = 0.f, te = 0.f;
            if (hostIO) cudaEventElapsedTime(&tin, ev[0], evPool[c * 6 + 5]);
            cudaEventElapsedTime(&tb, ev[0], evPool[c * 6 + 0]);
            cudaEventElapsedTime(&te, ev[0], evPool[c * 6 + 4]);
            std::fprintf(stderr, "[cpq] chunk %zu: h2d_done %.2f compute %.2f..%.2f ms\n", c, tin, tb, te);
        }
        std::fprintf(stderr, "[cpq] total %.2f ms, h2d span %.2f, d2h span %.2f\n", timings.total_ms, timings.h2d_ms, timings.d2h_ms);
    }
    if (tf[1] != 0)
    {
        CPQ_CUDA(cudaMemset(ticketFault.p + 1, 0, sizeof(unsigned)));
        setError("EQ filter state became non-finite or exceeded 1e15 where the engine had not foreseen it (look-back links with a start "
                 "state above 1e9): the reference resets the state there (EQProcessor.Processing.cpp:174-175)");
        return CPQ_ERR_UNSUPPORTED;
    }
    return CPQ_OK;
}

} // namespace cpq

// ================================================================================================
// C ABI
// ================================================================================================
using cpq::Engine;
struct cpq_engine : Engine {};

extern "C" {

int cpq_abi_version(void) { return CPQ_ABI_VERSION; }

int cpq_debug_check_guards(void) { return cpq::checkGuards(); }

const char* cpq_status_string(cpq_status s)
{
    switch (s)
    {
        case CPQ_OK: return "ok";
        case CPQ_ERR_INVALID: return "invalid argument";
        case CPQ_ERR_NOT_READY: return "not ready";
        case CPQ_ERR_CUDA: return "CUDA error / no device";
        case CPQ_ERR_OOM: return "out of memory";
        case CPQ_ERR_UNSUPPORTED: return "unsupported";
        case CPQ_ERR_GEOMETRY: return "layer geometry mismatch";
    }
    return "?";
}

static thread_local std::string g_createError;

const char* cpq_last_error(cpq_handle h) { return h ? h->err.c_str() : g_createError.c_str(); }

void cpq_filter_spec_default(cpq_filter_spec* out)
{
    if (out) cpq::filterSpecDefault(out);
}

void cpq_config_default(cpq_config* out)
{
    if (!out) return;
    std::memset(out, 0, sizeof(*out));
    out->device = 0;
    out->n_streams = 1;
    out->n_channels = 2;
    out->block_size = 512;
    out->sample_rate = 48000.0;
    out->max_samples = 48000 * 10 / 512 * 512;
    out->conv_boundary = CPQ_CONV_INNER;
    out->workspace_bytes = 0;
}

cpq_status cpq_create(const cpq_config* cfg, cpq_handle* out)
{
    if (!cfg || !out) return CPQ_ERR_INVALID;
    *out = nullptr;
    auto* e = new (std::nothrow) cpq_engine();
    if (!e) return CPQ_ERR_OOM;
    cpq_status st = e->init(cfg);
    if (st != CPQ_OK)
    {
        g_createError = e->err;
        delete e;
        return st;
    }
    *out = e;
    return CPQ_OK;
}

void cpq_destroy(cpq_handle h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    delete h;
}

cpq_status cpq_reset(cpq_handle h)
{
    if (!h) return CPQ_ERR_INVALID;
    return h->resetState();
}

cpq_status cpq_set_streaming(cpq_handle h, int enable)
{
    if (!h) return CPQ_ERR_INVALID;
    h->streaming = enable != 0;
    return h->resetState();
}

int64_t cpq_stream_position(cpq_handle h) { return h ? h->absCallback * (int64_t) h->cfg.block_size : -1; }

size_t cpq_state_size(cpq_handle h)
{
    if (!h || !h->streaming) return 0;
    if (h->ensureStreamState() != CPQ_OK) return 0;
    return h->stateBytes();
}

cpq_status cpq_export_state(cpq_handle h, void* dst, size_t bytes)
{
    if (!h) return CPQ_ERR_INVALID;
    return h->exportState(dst, bytes);
}

cpq_status cpq_import_state(cpq_handle h, const void* src, size_t bytes)
{
    if (!h) return CPQ_ERR_INVALID;
    return h->importState(src, bytes);
}

cpq_status cpq_set_impulse(cpq_handle h, int stream, int channel, const double* ir, int ir_len, double scale,
                           const cpq_filter_spec* spec)
{
    if (!h) return CPQ_ERR_INVALID;
    return h->setImpulse(stream, channel, ir, ir_len, scale, spec);
}

cpq_status cpq_set_eq(cpq_handle h, int stream, const cpq_svf_coeffs coeffs[CPQ_NUM_BANDS],
                      const uint8_t active[CPQ_NUM_BANDS], const int32_t chan_mode[CPQ_NUM_BANDS],
                      double saturation, double total_gain_lin)
{
    if (!h || !coeffs || !active || !chan_mode) return CPQ_ERR_INVALID;
    if (h->cfg.shared_eq ? (stream != -1 && stream != 0) : (stream < 0 || stream >= h->cfg.n_streams))
    {
        h->setError("set_eq: stream out of range");
        return CPQ_ERR_INVALID;
    }
    for (int b = 0; b < CPQ_NUM_BANDS; ++b)
    {
        if (active[b] && (chan_mode[b] < 0 || chan_mode[b] > 4)) return CPQ_ERR_INVALID;
        if (active[b] && chan_mode[b] >= 3 && h->cfg.n_channels != 2)
        {
            h->setError("set_eq: Mid/Side bands need a stereo handle (n_channels = 2)");
            return CPQ_ERR_UNSUPPORTED;
        }
    }
    if (!(saturation >= 0.0) || !std::isfinite(total_gain_lin)) return CPQ_ERR_INVALID;
    cpq::EqSet& e = h->eqSets[h->cfg.shared_eq ? 0 : (size_t) stream];
    std::memcpy(e.coeffs, coeffs, sizeof(e.coeffs));
    std::memcpy(e.active, active, sizeof(e.active));
    std::memcpy(e.mode, chan_mode, sizeof(e.mode));
    e.saturation = saturation;
    e.totalGain = total_gain_lin;
    e.set = true;
    e.events.clear();
    e.ramp = cpq::GainRampState {};   // the total gain is set, not ramped to
    h->eqDirty = true;
    return CPQ_OK;
}

cpq_status cpq_set_eq_mode(cpq_handle h, int stream, int filter_structure, int agc_enabled, const uint8_t node_active[CPQ_NUM_BANDS])
{
    if (!h || filter_structure < 0 || filter_structure > 1) return CPQ_ERR_INVALID;
    if (h->cfg.shared_eq ? (stream != -1 && stream != 0) : (stream < 0 || stream >= h->cfg.n_streams))
    {
        h->setError("set_eq_mode: stream out of range");
        return CPQ_ERR_INVALID;
    }
    cpq::EqSet& e = h->eqSets[h->cfg.shared_eq ? 0 : (size_t) stream];
    e.structure = filter_structure;
    e.agc = agc_enabled ? 1 : 0;
    e.hasNodeActive = node_active != nullptr;
    if (node_active) std::memcpy(e.nodeActive, node_active, sizeof(e.nodeActive));
    h->eqDirty = true;
    return CPQ_OK;
}

cpq_status cpq_get_agc_state(cpq_handle h, int stream, double out[3])
{
    if (!h || !out || stream < 0 || stream >= h->cfg.n_streams) return CPQ_ERR_INVALID;
    if (!h->anyAgc || !h->agcState.p) return CPQ_ERR_NOT_READY;
    Engine* e = h;
    auto setError = [&](const std::string& s) { e->setError(s); };
    CPQ_CUDA(cudaSetDevice(e->cfg.device));
    CPQ_CUDA(cudaMemcpy(out, e->agcState.p + (size_t) stream * 3, 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return CPQ_OK;
}

int cpq_band_node_active(int type, float gain_db, int enabled, double sample_rate)
{
    // createBandNode, EQProcessor.Coefficients.cpp:27-58
    if (!enabled || !(sample_rate > 0.0)) return 0;
    if (type != 3 && type != 4 && std::fabs(gain_db) < 0.01f) return 0;
    return 1;
}

cpq_status cpq_schedule_total_gain(cpq_handle h, int stream, int64_t at_callback, double new_gain_lin)
{
    if (!h || at_callback < 0 || !std::isfinite(new_gain_lin)) return CPQ_ERR_INVALID;
    if (h->cfg.shared_eq ? (stream != -1 && stream != 0) : (stream < 0 || stream >= h->cfg.n_streams)) return CPQ_ERR_INVALID;
    cpq::EqSet& e = h->eqSets[h->cfg.shared_eq ? 0 : (size_t) stream];
    e.events.push_back({ at_callback, new_gain_lin });
    h->gainTabCallbacks = -1;
    return CPQ_OK;
}

cpq_status cpq_set_epilogue(cpq_handle h, double makeup_gain, int dither_bits)
{
    if (!h || !std::isfinite(makeup_gain) || dither_bits > 32) return CPQ_ERR_INVALID;
    h->makeup = makeup_gain;
    h->ditherBits = dither_bits < 0 ? 0 : dither_bits;
    return CPQ_OK;
}

cpq_status cpq_set_output_filter(cpq_handle h, int enabled, int conv_is_last, int hc_mode, int lc_mode, int lp_mode)
{
    if (!h || hc_mode < 0 || hc_mode > 2 || lc_mode < 0 || lc_mode > 1 || lp_mode < 0 || lp_mode > 2) return CPQ_ERR_INVALID;
    h->outCfg.filterEnabled = enabled ? 1 : 0;
    h->outCfg.convIsLast = conv_is_last ? 1 : 0;
    h->outCfg.hc = hc_mode;
    h->outCfg.lc = lc_mode;
    h->outCfg.lp = lp_mode;
    h->postDirty = true;
    return CPQ_OK;
}

cpq_status cpq_ir_scale_factor(const double* ir_l, const double* ir_r, int len, const double* cur_l, const double* cur_r, int cur_len,
                               double cur_scale, cpq_ir_scale* out)
{
    if (!out || len < 0 || (len > 0 && !ir_l)) return CPQ_ERR_INVALID;
    const double* ch[2] = { ir_l, ir_r };
    const double* cur[2] = { cur_l, cur_r };
    cpq::irScaleFactor(ch, ir_l ? (ir_r ? 2 : 1) : 0, len, cur_l ? cur : nullptr, cur_l ? (cur_r ? 2 : 1) : 0, cur_len, cur_scale, out);
    return CPQ_OK;
}

cpq_status cpq_ir_min_phase(const double* ir, int len, double* out)
{
    if (!ir || !out || len <= 0) return CPQ_ERR_INVALID;
    return cpq::irMinimumPhase(ir, len, out) ? CPQ_OK : CPQ_ERR_UNSUPPORTED;
}

int cpq_ir_target_length(double sample_rate, double target_seconds) { return cpq::irTargetLength(sample_rate, target_seconds); }

int cpq_ir_prepare(const double* ir, int len, double sample_rate, double target_seconds, double* out, int out_capacity)
{
    if (!ir || !out || len <= 0 || !(sample_rate > 0.0)) return -1;
    const int target = cpq::irTargetLength(sample_rate, target_seconds);
    if (out_capacity < target) return -1;
    return cpq::irPrepare(ir, len, sample_rate, target_seconds, out);
}

double cpq_ir_freq_peak_gain(const double* ir_l, const double* ir_r, int len)
{
    if (!ir_l || len <= 0) return 1.0;
    const double* ch[2] = { ir_l, ir_r };
    return cpq::irFreqPeakGain(ch, ir_r ? 2 : 1, len);
}

int cpq_parse_eq_preset(const char* text, cpq_eq_band_params bands[CPQ_NUM_BANDS], float* total_gain_db)
{
    if (!text || !bands || !total_gain_db) return -1;
    return cpq::parseEqPreset(text, bands, total_gain_db);
}

cpq_status cpq_set_direct_head(cpq_handle h, int enable)
{
    if (!h) return CPQ_ERR_INVALID;
    if (h->planSet && (enable != 0) != (h->directHead != 0))
    {
        h->setError("set_direct_head: call before the first cpq_set_impulse (the head taps are removed from the partitions at SetImpulse time)");
        return CPQ_ERR_INVALID;
    }
    h->directHead = enable ? 1 : 0;
    return CPQ_OK;
}

cpq_status cpq_set_mix(cpq_handle h, float mix, int dry_delay_samples)
{
    if (!h || !(mix >= 0.0f && mix <= 1.0f) || dry_delay_samples < 0 || dry_delay_samples > 2097152 + 524288) return CPQ_ERR_INVALID;
    h->mix = static_cast<double>(mix);
    h->dryDelay = dry_delay_samples;
    return CPQ_OK;
}

cpq_status cpq_set_partial_sources(cpq_handle h, int n, const double* const* device_ptrs)
{
    if (!h || n < 0 || n > CPQ_MAX_PEERS || (n > 0 && !device_ptrs)) return CPQ_ERR_INVALID;
    for (int p = 0; p < n; ++p)
        if (!device_ptrs[p]) return CPQ_ERR_INVALID;
    h->nPeers = n;
    for (int p = 0; p < n; ++p) h->peers[p] = device_ptrs[p];
    return CPQ_OK;
}

cpq_status cpq_set_stream_window(cpq_handle h, int first_stream, int n_streams)
{
    if (!h || first_stream < 0 || first_stream >= h->cfg.n_streams || n_streams == 0 || n_streams < -1) return CPQ_ERR_INVALID;
    if (n_streams > 0 && first_stream + n_streams > h->cfg.n_streams) return CPQ_ERR_INVALID;
    h->winFirst = first_stream;
    h->winCount = n_streams;
    return CPQ_OK;
}

cpq_status cpq_set_input_gain(cpq_handle h, double gain)
{
    if (!h || !std::isfinite(gain)) return CPQ_ERR_INVALID;
    h->inputGain = gain;
    return CPQ_OK;
}

cpq_status cpq_set_peak_limiter(cpq_handle h, double release_ms)
{
    if (!h || !(release_ms >= 0.0) || !std::isfinite(release_ms)) return CPQ_ERR_INVALID;
    h->limiterMs = release_ms;
    return CPQ_OK;
}

cpq_status cpq_set_convolver_bypass(cpq_handle h, int bypassed)
{
    if (!h) return CPQ_ERR_INVALID;
    h->convBypassed = bypassed ? 1 : 0;
    return CPQ_OK;
}

int cpq_ir_peak_latency(const double* ir_l, const double* ir_r, int len)
{
    // estimatePeakLatencySamples, convolver/ConvolverProcessor.LoaderThread.cpp:149-207
    if (len <= 0 || (!ir_l && !ir_r)) return 0;
    double maxCentroid = 0.0;
    for (const double* d : { ir_l, ir_r })
    {
        if (!d) continue;
        double total = 0.0;
        for (int i = 0; i < len; ++i) total += d[i] * d[i];
        if (total < 1e-12) continue;
        double cum = 0.0;
        int cutoff = len - 1;
        for (int i = 0; i < len; ++i)
        {
            cum += d[i] * d[i];
            if (cum >= total * 0.999) { cutoff = i; break; }
        }
        double sumE = 0.0, sumW = 0.0;
        for (int i = 0; i <= cutoff; ++i)
        {
            const double e = d[i] * d[i];
            sumE += e;
            sumW += static_cast<double>(i) * e;
        }
        maxCentroid = std::max(maxCentroid, sumE > 0.0 ? sumW / sumE : 0.0);
    }
    return std::clamp(static_cast<int>(std::floor(maxCentroid + 0.5)), 0, len - 1);
}

cpq_status cpq_ir_decode_wav(const void* bytes, size_t n, cpq_ir_file* info, double* out, size_t out_capacity)
{
    if (!bytes || !info) return CPQ_ERR_INVALID;
    cpq::WavData w;
    std::string err;
    if (!cpq::irDecodeWav(static_cast<const uint8_t*>(bytes), n, w, err)) return CPQ_ERR_UNSUPPORTED;
    const size_t frames = w.ch.empty() ? 0 : w.ch[0].size();
    info->channels = w.channels;
    info->frames = (int64_t) frames;
    info->sample_rate = w.sampleRate;
    info->bits_per_sample = w.bitsPerSample;
    info->is_float = w.isFloat;
    if (out)
    {
        if (out_capacity < frames * (size_t) w.channels) return CPQ_ERR_INVALID;
        for (int c = 0; c < w.channels; ++c) std::copy(w.ch[(size_t) c].begin(), w.ch[(size_t) c].end(), out + (size_t) c * frames);
    }
    return CPQ_OK;
}

int cpq_ir_trim_silence(const double* ch0, const double* ch1, int n)
{
    return (ch0 && n > 0) ? cpq::irTrimTrailingSilence(ch0, ch1, n) : -1;
}

cpq_status cpq_ir_mixed_phase(const double* linear, const double* minimum, int len, double sample_rate, double lo_hz, double hi_hz, double* out)
{
    if (!linear || !minimum || !out || len <= 0) return CPQ_ERR_INVALID;
    return cpq::irMixedPhaseFallback(linear, minimum, len, sample_rate, lo_hz, hi_hz, out) ? CPQ_OK : CPQ_ERR_UNSUPPORTED;
}

// LoaderThread::doLoadStep -> doTrimStep -> doTransformStep -> doBuildStep (convolver/ConvolverProcessor.LoaderThread.cpp:430-757)
// for one stream, with the pieces above
cpq_status cpq_load_impulse_wav(cpq_handle h, int stream, const void* bytes, size_t n, int phase_mode, double target_seconds,
                                const cpq_filter_spec* spec, cpq_ir_load_info* info)
{
    if (!h || !bytes || phase_mode < 0 || phase_mode > 2 || !(target_seconds > 0.0)) return CPQ_ERR_INVALID;
    Engine* e = h;
    cpq::WavData w;
    std::string err;
    if (!cpq::irDecodeWav(static_cast<const uint8_t*>(bytes), n, w, err))
    {
        e->setError("load impulse: " + err);
        return CPQ_ERR_UNSUPPORTED;
    }
    const int frames = (int) w.ch[0].size();
    if (frames <= 0)
    {
        e->setError("load impulse: empty file");
        return CPQ_ERR_INVALID;
    }
    const int nch = std::min(w.channels, 2);
    const int len = cpq::irTrimTrailingSilence(w.ch[0].data(), nch > 1 ? w.ch[1].data() : nullptr, frames);
    if (std::fabs(w.sampleRate - e->cfg.sample_rate) > 1e-6)
    {
        e->setError("load impulse: the file's sample rate differs from the engine's; the reference resamples with r8brain-free-src "
                    "(third party), which this library does not contain: resample first");
        return CPQ_ERR_UNSUPPORTED;
    }
    const int target = cpq::irTargetLength(w.sampleRate, target_seconds);
    std::vector<std::vector<double>> ir((size_t) nch, std::vector<double>((size_t) target));
    for (int c = 0; c < nch; ++c) cpq::irPrepare(w.ch[(size_t) c].data(), len, w.sampleRate, target_seconds, ir[(size_t) c].data());
    // validateBuffer (:645-659): finite and not silent
    auto valid = [&](const std::vector<std::vector<double>>& b) {
        double mx = 0.0;
        for (auto& v : b)
            for (double x : v)
            {
                if (!std::isfinite(x)) return false;
                mx = std::max(mx, std::fabs(x));
            }
        return mx > 1.0e-12;
    };
    int applied = 0;
    if (phase_mode >= 1)
    {
        std::vector<std::vector<double>> mp((size_t) nch, std::vector<double>((size_t) target));
        bool ok = true;
        for (int c = 0; c < nch && ok; ++c) ok = cpq::irMinimumPhase(ir[(size_t) c].data(), target, mp[(size_t) c].data());
        if (ok && valid(mp))
        {
            if (phase_mode == 1)
            {
                ir = mp;
                applied = 1;
            }
            else
            {
                std::vector<std::vector<double>> mx((size_t) nch, std::vector<double>((size_t) target));
                bool ok2 = true;
                for (int c = 0; c < nch && ok2; ++c)
                    ok2 = cpq::irMixedPhaseFallback(ir[(size_t) c].data(), mp[(size_t) c].data(), target, w.sampleRate, 200.0, 1000.0, mx[(size_t) c].data());
                if (ok2 && valid(mx))
                {
                    ir = mx;
                    applied = 2;
                }
            }
        }
    }
    const double* chp[2] = { ir[0].data(), nch > 1 ? ir[1].data() : nullptr };
    cpq_ir_scale sc {};
    cpq::irScaleFactor(chp, nch, target, nullptr, 0, 0, 1.0, &sc);
    const double scale = sc.has_scale_factor ? sc.scale_factor : 1.0;
    const int peak = cpq_ir_peak_latency(chp[0], chp[1], target);
    for (int c = 0; c < e->cfg.n_channels; ++c)
    {
        const cpq_status st = e->setImpulse(stream, c, (c == 1 && nch > 1) ? chp[1] : chp[0], target, scale, spec);
        if (st != CPQ_OK) return st;
    }
    if (info)
    {
        info->file_channels = w.channels;
        info->file_frames = frames;
        info->file_sample_rate = w.sampleRate;
        info->trimmed_frames = len;
        info->target_length = target;
        info->peak_latency = peak;
        info->scale_factor = scale;
        info->phase_applied = applied;
    }
    return CPQ_OK;
}

cpq_status cpq_set_conv_input_trim(cpq_handle h, double gain)
{
    if (!h || !std::isfinite(gain)) return CPQ_ERR_INVALID;
    h->convInputTrim = gain;
    return CPQ_OK;
}

cpq_status cpq_set_output_stage(cpq_handle h, double dc_cutoff_hz, int hard_clamp)
{
    if (!h || !(dc_cutoff_hz >= 0.0) || !std::isfinite(dc_cutoff_hz)) return CPQ_ERR_INVALID;
    h->outCfg.dcCutoff = dc_cutoff_hz;
    h->outCfg.finalClamp = hard_clamp ? 1 : 0;
    h->postDirty = true;
    return CPQ_OK;
}

void cpq_output_filter_design(double sample_rate, int conv_is_last, int hc_mode, int lc_mode, int lp_mode, double out[15])
{
    cpq::BiquadC st[3];
    cpq::outputFilterStages(sample_rate, conv_is_last, hc_mode, lc_mode, lp_mode, st);
    for (int i = 0; i < 3; ++i)
    {
        out[5 * i] = st[i].b0; out[5 * i + 1] = st[i].b1; out[5 * i + 2] = st[i].b2; out[5 * i + 3] = st[i].a1; out[5 * i + 4] = st[i].a2;
    }
}

cpq_status cpq_set_dither_uniforms(cpq_handle h, const double* uniforms, int64_t samples_per_channel)
{
    if (!h || !uniforms || samples_per_channel <= 0) return CPQ_ERR_INVALID;
    Engine* e = h;
    auto setError = [&](const std::string& s) { e->setError(s); };
    CPQ_CUDA(cudaSetDevice(e->cfg.device));
    const size_t n = (size_t) e->nSeq * 2 * (size_t) samples_per_channel;
    CPQ_CUDA(e->uniforms.ensure(n));
    CPQ_CUDA(cudaMemcpy(e->uniforms.p, uniforms, n * sizeof(double), cudaMemcpyHostToDevice));
    e->uniformsPerCh = samples_per_channel;
    e->uniformsBorrowed = nullptr;
    e->ditherRng = false;
    return CPQ_OK;
}

cpq_status cpq_set_dither_seed(cpq_handle h, const uint64_t* stream_seeds)
{
    if (!h) return CPQ_ERR_INVALID;
    Engine* e = h;
    auto setError = [&](const std::string& s) { e->setError(s); };
    if (!stream_seeds)
    {
        e->ditherRng = false;
        return CPQ_OK;
    }
    CPQ_CUDA(cudaSetDevice(e->cfg.device));
    // PsychoacousticDither(seed): SplitMix64(seed) hands every channel i its seedValue; fallbackState[i] = seedValue ^ 0xd1b5...
    // (PsychoacousticDither.h:118-140); channel c of a stream is channel c of that stream's own dither object
    e->rngSeedState.assign((size_t) e->nSeq, 0ull);
    for (int st = 0; st < e->cfg.n_streams; ++st)
    {
        unsigned long long sm = stream_seeds[st];
        for (int c = 0; c < e->cfg.n_channels; ++c)
        {
            unsigned long long z = (sm += 0x9e3779b97f4a7c15ull);
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            z ^= z >> 31;
            e->rngSeedState[(size_t) st * e->cfg.n_channels + c] = z ^ 0xd1b54a32d192ed03ull;
        }
    }
    CPQ_CUDA(e->rngState.ensure((size_t) e->nSeq));
    CPQ_CUDA(cudaMemcpy(e->rngState.p, e->rngSeedState.data(), (size_t) e->nSeq * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    e->ditherRng = true;
    return CPQ_OK;
}

cpq_status cpq_set_dither_uniforms_device(cpq_handle h, const double* d_uniforms, int64_t samples_per_channel)
{
    if (!h || !d_uniforms || samples_per_channel <= 0 || (reinterpret_cast<uintptr_t>(d_uniforms) & 15)) return CPQ_ERR_INVALID;
    h->uniformsBorrowed = d_uniforms;
    h->uniformsPerCh = samples_per_channel;
    h->ditherRng = false;
    return CPQ_OK;
}

cpq_status cpq_design_band(int type, float freq_hz, float gain_db, float q, double sample_rate, cpq_svf_coeffs* out)
{
    if (!out) return CPQ_ERR_INVALID;
    return cpq::designBand(type, freq_hz, gain_db, q, sample_rate, out) ? CPQ_OK : CPQ_ERR_INVALID;
}
double cpq_db_to_gain(float db) { return cpq::dbToGain(db); }
double cpq_equal_power_sin(double x) { return cpq::equalPowerSin(x); }

cpq_status cpq_process_device(cpq_handle h, double* d_io, int64_t stride, int64_t T, unsigned stages)
{
    if (!h) return CPQ_ERR_INVALID;
    return h->processDevice(d_io, stride, T, stages);
}

cpq_status cpq_process(cpq_handle h, double* const* planar, int64_t T, unsigned stages)
{
    if (!h || !planar) return CPQ_ERR_INVALID;
    Engine* e = h;
    auto setError = [&](const std::string& s) { e->setError(s); };
    if (T <= 0 || T > e->cfg.max_samples || T % e->cfg.block_size != 0)
    {
        e->setError("process: T must be a positive multiple of block_size and <= max_samples");
        return CPQ_ERR_INVALID;
    }
    for (int s = 0; s < e->nSeq; ++s)
        if (!planar[s])
        {
            e->setError("process: null channel pointer");
            return CPQ_ERR_INVALID;
        }
    CPQ_CUDA(cudaSetDevice(e->cfg.device));
    const int64_t stride = (T + 1) & ~(int64_t) 1;
    CPQ_CUDA(e->io.ensure((size_t) e->nSeq * stride));
    // H2D, kernels and D2H are pipelined per sequence chunk on three streams inside processCore
    return e->processCore(e->io.p, stride, T, stages, planar);
}

cpq_status cpq_process_f32(cpq_handle h, float* const* planar, int64_t T, unsigned stages)
{
    if (!h || !planar) return CPQ_ERR_INVALID;
    Engine* e = h;
    auto setError = [&](const std::string& s) { e->setError(s); };
    if (T <= 0 || T > e->cfg.max_samples || T % e->cfg.block_size != 0)
    {
        e->setError("process: T must be a positive multiple of block_size and <= max_samples");
        return CPQ_ERR_INVALID;
    }
    for (int s = 0; s < e->nSeq; ++s)
        if (!planar[s])
        {
            e->setError("process: null channel pointer");
            return CPQ_ERR_INVALID;
        }
    CPQ_CUDA(cudaSetDevice(e->cfg.device));
    const int64_t stride = (T + 1) & ~(int64_t) 1;
    CPQ_CUDA(e->io.ensure((size_t) e->nSeq * stride));
    return e->processCore(e->io.p, stride, T, stages, nullptr, planar);
}

cpq_status cpq_set_partition_range(cpq_handle h, int part_begin, int part_end)
{
    if (!h || part_begin < 0) return CPQ_ERR_INVALID;
    h->partBegin = part_begin;
    h->partEnd = part_end;
    return CPQ_OK;
}

int cpq_total_partitions(cpq_handle h)
{
    if (!h || !h->planSet) return 0;
    int total = 0;
    for (int li = 0; li < h->plan.numLayers; ++li) total += h->plan.layers[li].numPartsIR;
    return total;
}

static void fillLayout(const cpq::ConvPlan& plan, const cpq::GatherPlan* g, cpq_layout* out)
{
    std::memset(out, 0, sizeof(*out));
    out->num_layers = plan.numLayers;
    out->latency = plan.numLayers > 0 ? plan.layers[0].partSize : 0;
    for (int li = 0; li < plan.numLayers; ++li)
    {
        const cpq::LayerPlan& l = plan.layers[li];
        cpq_layer_layout& o = out->layers[li];
        o.part_size = l.partSize;
        o.fft_size = l.fftSize;
        o.num_parts_ir = l.numPartsIR;
        o.num_parts = l.numParts;
        o.parts_per_callback = l.partsPerCallback;
        o.output_delay_samples = l.outputDelaySamples;
        o.ir_offset = l.irOffset;
        o.ir_len = l.irLen;
        o.gain = l.gain;
        o.first_output_sample = g ? g->firstOutput[li] : -1;
        o.skipped_callbacks = g ? g->skipped[li] : 0;
    }
}

cpq_status cpq_get_layout(cpq_handle h, cpq_layout* out)
{
    if (!h || !out) return CPQ_ERR_INVALID;
    if (!h->planSet) return CPQ_ERR_NOT_READY;
    cpq::GatherPlan g;
    cpq::simulateCallbacks(h->plan, h->cfg.max_samples / h->cfg.block_size, g);
    fillLayout(h->plan, &g, out);
    return CPQ_OK;
}

int cpq_latency(cpq_handle h) { return (h && h->planSet) ? h->plan.layers[0].partSize : 0; }

cpq_status cpq_get_timings(cpq_handle h, cpq_timings* out)
{
    if (!h || !out) return CPQ_ERR_INVALID;
    *out = h->timings;
    return CPQ_OK;
}

cpq_status cpq_get_eq_state(cpq_handle h, int stream, double* out)
{
    if (!h || !out || stream < 0 || stream >= h->cfg.n_streams) return CPQ_ERR_INVALID;
    Engine* e = h;
    auto setError = [&](const std::string& s) { e->setError(s); };
    CPQ_CUDA(cudaSetDevice(e->cfg.device));
    const size_t per = (size_t) e->cfg.n_channels * CPQ_NUM_BANDS * 2;
    CPQ_CUDA(cudaMemcpy(out, e->stateOut.p + (size_t) stream * per, per * sizeof(double), cudaMemcpyDeviceToHost));
    return CPQ_OK;
}

void* cpq_cuda_stream(cpq_handle h) { return h ? (void*) h->stream : nullptr; }
int64_t cpq_kernel_launch_count(cpq_handle h) { return h ? h->launches : 0; }

cpq_status cpq_plan_layout(int ir_len, int block_size, const cpq_filter_spec* spec, int64_t n_callbacks,
                           cpq_layout* out, int64_t* src_offsets)
{
    if (!out || n_callbacks < 0) return CPQ_ERR_INVALID;
    cpq::ConvPlan plan;
    if (!cpq::makeConvPlan(ir_len, block_size, spec, plan)) return CPQ_ERR_INVALID;
    cpq::GatherPlan g;
    cpq::simulateCallbacks(plan, n_callbacks, g);
    fillLayout(plan, &g, out);
    if (src_offsets)
        for (int li = 1; li < plan.numLayers; ++li)
            std::memcpy(src_offsets + (size_t) (li - 1) * n_callbacks, g.tailSrc[li].data(), (size_t) n_callbacks * sizeof(int64_t));
    return CPQ_OK;
}

cpq_status cpq_plan_layout_ex(int ir_len, int known_block_size, int call_size, const cpq_filter_spec* spec, int64_t n_callbacks,
                              cpq_layout* out, int64_t* src_offsets, int64_t* l0_src, int32_t* l0_count)
{
    if (!out || n_callbacks < 0 || call_size <= 0) return CPQ_ERR_INVALID;
    cpq::ConvPlan plan;
    if (!cpq::makeConvPlan(ir_len, known_block_size, spec, plan)) return CPQ_ERR_INVALID;
    plan.callSize = call_size;
    cpq::GatherPlan g;
    cpq::simulateCallbacks(plan, n_callbacks, g);
    fillLayout(plan, &g, out);
    if (src_offsets)
        for (int li = 1; li < plan.numLayers; ++li)
            std::memcpy(src_offsets + (size_t) (li - 1) * n_callbacks, g.tailSrc[li].data(), (size_t) n_callbacks * sizeof(int64_t));
    for (int64_t c = 0; c < n_callbacks; ++c)
    {
        if (l0_src) l0_src[c] = g.l0Identity ? c * (int64_t) call_size : g.l0Src[(size_t) c];
        if (l0_count) l0_count[c] = g.l0Identity ? call_size : g.l0Count[(size_t) c];
    }
    return CPQ_OK;
}

/* Dependent DFMA latency in SM cycles (one warp, one chain). */
double cpq_probe_dfma_latency(int device)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    double* d = nullptr;
    long long* c = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess || cudaMalloc(&c, 8) != cudaSuccess) return -1.0;
    const int iters = 1 << 16;
    cpq::dfma_latency_kernel<<<1, 32>>>(d, c, iters);
    cpq::dfma_latency_kernel<<<1, 32>>>(d, c, iters);
    long long cyc = 0;
    cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaFree(c);
    return cudaGetLastError() == cudaSuccess ? (double) cyc / iters : -1.0;
}

/* DFMA throughput probe used by bench.py for the FP64-pipe roofline (not part of the reference path). */
double cpq_probe_dfma_tflops(int device, int iters)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop {};
    cudaGetDeviceProperties(&prop, device);
    double* d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    cpq::dfma_probe_kernel<<<blocks, threads>>>(d, 64);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    cpq::dfma_probe_kernel<<<blocks, threads>>>(d, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess || ms <= 0.f) return -1.0;
    const double flops = 2.0 * 8.0 * (double) iters * (double) blocks * threads;
    return flops / (ms * 1e-3) / 1e12;
}
}
