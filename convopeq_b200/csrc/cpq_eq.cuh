// cpq_eq.cuh -- layer assembly + 20-band TPT-SVF EQ + gain/headroom epilogue, and the dither kernel (sm_100a).
//
//   eq_kernel      Get()'s layer sum (MKLNonUniformConvolver.cpp:1553-1634) -> optional ConvolverProcessor wet
//                  scrub/gain (ConvolverProcessor.Runtime.cpp:50-60,675,748) -> 20 x processBandStereo
//                  (EQProcessor.Processing.cpp:191-276) -> total-gain ramp (:1262-1274) -> makeup gain +
//                  kOutputHeadroom (AudioEngine.Processing.DSPCoreDouble.cpp:465-469,:655-663)
//   dither_kernel  PsychoacousticDither::processStereoBlock (PsychoacousticDither.h:293-355)
//
// The band recurrence is linear in its 2-element state (the tanh saturation only touches the band *output*,
// Processing.cpp:148-168), so each band is a blocked scan.  A CTA owns one 8192-sample tile of one sequence, split
// into eight 1024-sample segments, one per warp, 32 consecutive samples per thread in registers.  Per band a warp
// runs: pass 1 = zero-state response of each thread's block by 32 precomputed weight vectors, warp-shuffle scan of
// s -> A^32 s + c, the *link* (below), a two-level table lookup for A^(32 lane), and pass 2 = the band recurrence
// from the exact start state.  Bands are processed in order on the same registers, so a tile crosses HBM once for
// all 20 bands.
//
// Links instead of barriers: the warps of a CTA are a chain.  Warp w takes the state at the start of its segment
// for band b from a shared-memory mailbox st[b][w] (flag fl[b][w]), and its lane 31 posts A^1024 s_in + (segment
// aggregate) into st[b][w+1] right after the warp scan.  There is no CTA barrier in the band loop: warps run
// skewed by one link latency and otherwise independently, and each warp loads/stores its own segment.  Tiles of a
// sequence are chained the same way through 16-byte records (seq, tile, band) in global memory whose payload is its
// own flag (the host fills them with all-ones before the launch): the last warp writes the state after the tile,
// the first warp of the next tile reads it -- one band ahead of its use, so the L2 round trip is hidden and all
// eight warps compute (a dedicated chain warp left one SM sub-partition with half the warps of the others:
// measured 67 % vs 77 % FP64-pipe ceiling for the band loops, scripts/probes/eq_inner_probe.cu).  Tiles take
// their index from an atomic ticket in run-major order, so a predecessor CTA is always already running (or done)
// when a CTA starts.
//
// Pass 2 arithmetic.  The reference designs every band as a TPT SVF (a2 = g a1, a3 = g a2,
// EQProcessor.Coefficients.cpp:101-130), which gives v2 = ic2 + g v1 and ic2' = ic2 + 2 g v1; with
// v1 = a1 ic1 - a2 ic2 + a2 v0 a Peaking band (m0 = 1, m2 = 0) costs 6 FP64 instructions per sample instead of the
// 8 of the literal form, other TPT bands 8 instead of 10.  Coefficient sets that are not TPT-consistent (raw
// cpq_set_eq input) run the literal form.  Differences to the reference's association are O(1e-16) relative.
//
// Fast path / exact path: with |out| < 4.5 before saturation the reference's clamps and scrubs are
// identities, so pass 2 runs without them and tracks a per-thread maximum; a flagged thread replays its block
// from the stashed inputs with the reference's full per-sample semantics.  Where a *state* can leave the
// reference's valid range (|ic| >= 1e15 or non-finite: the reference zeroes it and carries on, :174-175/:257-258) the
// scan's linearity assumption is void: a warp whose segment may see such a reset for a band (non-finite or huge samples
// in the band's input, a start state above 1e9, raw coefficient sets outside the TPT contract) runs that band over its
// 1024 samples lane after lane with the literal recurrence, resets included, and posts the true state after the
// segment to its successor instead of the scan's; everything downstream continues on the fast path (serialBand).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>

#include "../../include/cpq.h"
#include "cpq_mac.cuh"   // mbarrier / bulk-copy helpers

namespace cpq
{

#ifndef CPQ_EQ_L
#define CPQ_EQ_L 32
#endif
#ifndef CPQ_MAX_PEERS
#define CPQ_MAX_PEERS 8
#endif
#ifndef CPQ_EQ_WARPS
#define CPQ_EQ_WARPS 8
#endif
constexpr int kEqCWarps = CPQ_EQ_WARPS;          // every warp computes (an idle chain warp would leave one SM sub-partition half empty)
constexpr int kEqCThreads = kEqCWarps * 32;      // 256
constexpr int kEqThreads = kEqCThreads;
static_assert(kEqCWarps >= 6 && kEqCWarps <= 8, "mailbox rows hold 8 warps; 24 stages x 8 flags are cleared by the first 192 threads");
constexpr int kEqL = CPQ_EQ_L;                   // samples per compute thread (16 or 32)
constexpr int kEqTile = kEqCThreads * kEqL;      // 4096 (8192)
constexpr int kEqPad = kEqL + 2;                 // shared-memory doubles per thread block (keeps 16-byte alignment)
static_assert(kEqL == 16 || kEqL == 32, "samples per thread");

// per (parameter set, band) constants, all double; built by buildBandConstants() in cpq_engine.cu
constexpr int kEqcCoef = 0;                      // a1,a2,a3,m0,m1,m2, forceExact(!=0), kind (0 literal, 1 TPT peaking, 2 TPT), g, 2g, 2(a1-a3), 2(a1+a3)
                                                 // kind 3 (DF2T biquad): b0,b1,b2,a1,a2 in [0..4]; kind 4 (DC blocker): alpha0, alpha1 in [0..1]
constexpr int kEqcW = 12;                        // w[L][2]    zero-state weights, c = sum_j w[j] * v0[j]
constexpr int kEqcMs = kEqcW + 2 * kEqL;         // Ms[5][4]   A^(L*2^d), row-major 2x2
constexpr int kEqcPlo = kEqcMs + 20;             // Plo[8][4]  A^(L*j)
constexpr int kEqcPhi = kEqcPlo + 32;            // Phi[4][4]  A^(8L*j)
constexpr int kEqcMw = kEqcPhi + 16;             // A^(32L)    (one warp segment)
constexpr int kEqcMh = kEqcMw + 4;               // A^(L/2)    (half a thread block: start state of the second pass-2 chain)
constexpr int kEqcPw = kEqcMh + 4;               // Pw[9][4]   A^(32L*j), j = 0..8: whole warp segments (look-back composition)
constexpr int kEqcStride = kEqcPw + 36;          // doubles per band, staged in shared memory for all 20 bands

constexpr int kEqSeg = 32 * kEqL;                // 512 samples per warp segment
// Stages of the scan pipeline: the 20 EQ bands, then the linear output stages of DSPCore::processDouble --
// OutputFilter's three DF2T biquads (OutputFilter.cpp:199-421) and the two-section output DC blocker
// (UltraHighRateDCBlocker.h:98-126) -- each a 2-state linear recurrence handled by the same scan machinery.
constexpr int kEqPostStages = 4;                 // 0..2 OutputFilter biquads, 3 DC blocker
constexpr int kEqStages = CPQ_NUM_BANDS + kEqPostStages;
constexpr int kEqStageDc = CPQ_NUM_BANDS + 3;
constexpr int kEqNoSerial = 0, kEqBailOut = 1, kEqSerial = 2;   // what a band does when a state reset is possible in its segment (bandStart)
// shared memory: segment tiles | band constants | mailboxes st[20][8] (double2) | flags fl[20][8] (int) | ticket
// shared memory: segment tiles | EQ band constants | mailboxes st[stage][8] (double2) | flags fl[stage][8] (int) | ticket |
// output-stage constants (only allocated when such a stage runs)
// mailboxes: agg[stage][8] warp aggregates + tin[stage] tile-in states (double2 each), one mbarrier per mailbox entry; the
// look-back form adds endSt[stage][8], the states after segments that ran serially, with their mbarriers
constexpr int kEqMail = kEqStages * 9;
constexpr int kEqMailLb = kEqStages * 8;
__host__ __device__ constexpr int eqSmemDoubles(bool lb)
{
    return kEqCThreads * kEqPad + CPQ_NUM_BANDS * (lb ? kEqcStride : kEqcPw) + kEqMail * 2 + kEqMail + 2 + (lb ? kEqMailLb * 3 : 0);
}
constexpr size_t eqSmemBytes(bool lb, bool post)
{
    return (size_t) (eqSmemDoubles(lb) + (post ? kEqPostStages * (lb ? kEqcStride : kEqcPw) : 0)) * sizeof(double);
}
constexpr int kEqSmemDoubles = eqSmemDoubles(true);
constexpr size_t kEqSmemBytes = (size_t) kEqSmemDoubles * sizeof(double);
constexpr size_t kEqSmemBytesPost = kEqSmemBytes + (size_t) kEqPostStages * kEqcStride * sizeof(double);

struct EqChain
{
    double2* rec;         // [(seq*nRuns + run)*20 + band] = band state after the tile; all-ones words = not written yet
    unsigned* ticket;     // CTA ticket counter
};

struct EqArgs
{
    double* io;             // [nSeq][ioStride] in/out (in place); holds y0 (or the raw input when !assemble)
    int64_t ioStride;
    int64_t T;              // samples per sequence
    int nSeq;
    int nRuns;              // tiles per sequence = ceil(T / kEqTile); grid = nSeq * nRuns
    // assembly
    int assemble;           // add tails / apply the outer boundary
    int nTail;              // number of tail layers (0..2)
    const double* tail[2];  // [nSeq][tailStride[l]] layer output streams
    int64_t tailStride[2];
    const int64_t* tailSrc[2];     // [nCallbacks] stream position or -1
    const int32_t* blockMap[2];    // nullable: stream block -> frame
    int tailPartLog2[2];
    double tailGain[2];
    int blockLog2;          // log2(block size), or -1 when the host block is not a power of two (then blockSize divides)
    int blockSize;          // samples per callback
    // L0 through the reference's output ring when the host block differs from the L0 partition (non-power-of-two hosts:
    // SetImpulse gets the block rounded up, Add/Get run with the host block): out[c B + i] = l0[l0Src[c] + i] for i < l0Count[c]
    const double* l0;       // nullable [nSeq][l0Stride]; then io holds the convolver input, not the L0 output
    int64_t l0Stride;
    const int64_t* l0Src;
    const int32_t* l0Count;
    int outer;              // CPQ_CONV_OUTER: scrub + wet gain
    double wetGain;
    // EQ
    int doEq;
    const double* eqc;      // [nSets][20][kEqcStride]
    const unsigned* bandMask;  // [nSeq] bit b = band b processed for this sequence
    const unsigned* scalarMask;// nullable [nSeq] bit b = the reference runs band b of this sequence through the scalar processBand
                               // (Left / Right / Mid / Side bands, mono streams) whose tanh returns +-1 outside +-4.5
    unsigned scalarAll;        // the same for every sequence of the launch (Mid / Side rows)
    const int* setOfSeq;    // [nSeq]
    const double* sat;      // [nSets]
    double* stateOut;       // [nSeq][20][2] final states
    const double* stateIn;  // nullable [nSeq][20][2]: states at the start of the call (streaming continuation); null = Reset
    const double* postStateIn;   // nullable [nSeq][kEqPostStages][2]: the same for the output stages
    // linear output stages (stage index 20..23); run when postMask != 0, also without the EQ bands
    const double* postc;    // [kEqPostStages][kEqcStride]
    unsigned postMask;      // bit i = post stage i enabled
    double* postStateOut;   // nullable [nSeq][kEqPostStages][2] final states of the output stages
    int finalClamp;         // after the headroom multiply: bit 0 scrub (non-finite or |x| >= 1e300 -> 0), bit 1 clamp to +-kOutputHeadroom
    unsigned* limFlag;      // nullable [nSeq / limDiv]: set when a stored sample could engage the peak limiter (|x| > threshold - knee/2)
    int limDiv;             // channels per stream
    const double* gainTab;  // nullable [rows][nCallbacks][2] (start, inc); row = parameter set, or stream when gainBySeq
    const double* gainConst;// [nSets] settled total gain (used when gainTab == nullptr)
    int64_t nCallbacks;
    int doGain;             // apply the total gain (Processing.cpp:1262-1274) / the AGC gain ramp (:441-444) in this launch
    int gainBySeq;          // > 0: gainTab row = seq / gainBySeq (per-stream AGC table built by agc_kernel)
    unsigned bandSelect;    // bands this launch may run (a band sequence split around Mid/Side bands runs in several launches)
    unsigned bandSelectPar; // the same for sequences in the Parallel structure (all their bands must share one launch)
    // AGC block statistics (calculateRMS, Processing.cpp:21-52): per-callback sums of squares of the EQ input / of the band
    // output before the gain, [nSeq][nCallbacks]; nullable
    double* sumsqIn;
    double* sumsqOut;
    // Partition-range sharding: the convolver partials of all ranks, summed in rank order while the tile is
    // loaded -- peer[p] is rank p's buffer in the layout of io (this rank's own buffer included), mapped into this process
    // over NVLink.  The reduce step of the collective is this load; there is no separate pass and no staging copy.
    int nPeers;                     // 0 = io only
    const double* peer[CPQ_MAX_PEERS];
    // epilogue
    int doEpilogue;
    double makeup;
    int applyHeadroom;      // 1: multiply by kOutputHeadroom (no-dither branch)
    unsigned* fault;        // set to 1 when a state left the linear regime
    EqChain chain;
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_cg_f64(const double* p)
{
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// 16-byte record whose payload is its own flag: each 8-byte half is written atomically and is valid iff it is not the
// all-ones sentinel the host fills the records with before every launch
__device__ __forceinline__ double2 ld_volatile_f64x2(const double2* p)
{
    double2 v;
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_f64x2(double2* p, double2 v)
{
    asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ bool rec_pending(double2 v) { return __double2hiint(v.x) == -1 || __double2hiint(v.y) == -1; }
__device__ __forceinline__ double rec_clean(double v) { return __double2hiint(v) == -1 ? __longlong_as_double(0x7ff8000000000000ll) : v; }
__device__ __forceinline__ int lds_volatile(const int* p)
{
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned) __cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile(int* p, int v)
{
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned) __cvta_generic_to_shared(p)), "r"(v) : "memory");
}

__device__ __forceinline__ bool eq_valid(double v) { return fabs(v) < 1.0e15; }   // false for NaN / Inf too

// padded shared index: 18-double stride per 16 samples keeps every thread's block 16-byte aligned and both the
// coalesced pass (consecutive t) and the per-thread LDS.128/STS.128 pass free of bank conflicts
__device__ __forceinline__ int eq_sidx(int t) { return t + 2 * (t / kEqL); }

// The reference's per-sample semantics (processBandStereo, EQProcessor.Processing.cpp:191-276) over one thread's block,
// in registers: literal recurrence, clamped tanh 27/9 with a true division, scrubs and the +-100 clamp.  Used by warps
// in exact mode only.  Returns true if a state had to be reset (which the scan cannot represent).
// scalarTanh: the band goes through processBand (:128-186), whose fastTanh returns +-1 outside +-4.5 (FastTanhApprox.h:101-107)
// where the SSE form clamps its argument and evaluates (:112-119) -- the two differ by 1.6 % of `sat` beyond the threshold.
// capAt > 0: (cap1, cap2) receive the state after capAt samples (the sequence's final state when it ends inside this block).
__device__ __forceinline__ bool eq_pass2_exact(double (&x)[kEqL], double& ic1, double& ic2, const double* __restrict__ bc, double sat,
                                               bool scalarTanh, int capAt = 0, double* cap1 = nullptr, double* cap2 = nullptr)
{
    const double a1 = bc[0], a2 = bc[1], a3 = bc[2], m0 = bc[3], m1 = bc[4], m2 = bc[5];
    bool reset = false;
    const double oneMinusSat = 1.0 - sat;
#pragma unroll
    for (int j = 0; j < kEqL; ++j)
    {
        const double v0 = x[j];
        const double v3 = v0 - ic2;
        const double v1 = fma(a1, ic1, a2 * v3);
        const double v2 = fma(a2, ic1, fma(a3, v3, ic2));
        ic1 = fma(2.0, v1, -ic1);
        ic2 = fma(2.0, v2, -ic2);
        double out = fma(m0, v0, fma(m1, v1, m2 * v2));
        if (sat > 0.0)
        {
            double xc = (out > -4.5) ? out : -4.5;   // _mm_max_pd(x, lo): NaN -> lo
            xc = (xc < 4.5) ? xc : 4.5;
            const double x2 = xc * xc;
            double th = (xc * (27.0 + x2)) / fma(9.0, x2, 27.0);
            if (scalarTanh)
            {
                if (out >= 4.5) th = 1.0;
                else if (out <= -4.5) th = -1.0;
                else if (out != out) th = out;   // neither branch taken: the polynomial of a NaN
            }
            out = out * oneMinusSat + th * sat;
        }
        if (!eq_valid(out)) out = 0.0;
        if (!eq_valid(ic1)) { ic1 = 0.0; reset = true; }
        if (!eq_valid(ic2)) { ic2 = 0.0; reset = true; }
        out = (out > -100.0) ? out : -100.0;
        out = (out < 100.0) ? out : 100.0;
        x[j] = out;
        if (j + 1 == capAt)
        {
            *cap1 = ic1;
            *cap2 = ic2;
        }
    }
    return reset;
}

__device__ __forceinline__ void matvec2(const double* __restrict__ m, double& p1, double& p2, double add1, double add2)
{
    const double2 r0 = *reinterpret_cast<const double2*>(m), r1 = *reinterpret_cast<const double2*>(m + 2);
    const double n1 = fma(r0.x, p1, fma(r0.y, p2, add1));
    const double n2 = fma(r1.x, p1, fma(r1.y, p2, add2));
    p1 = n1;
    p2 = n2;
}

// Pass 2 (fast path) over one thread's block.
//   KIND 1  TPT Peaking  v1 = a1 ic1 - a2 ic2 + a2 v0;  out = v0 + m1 v1;  ic1' = 2 v1 - ic1;  ic2' = ic2 + 2g v1
//   KIND 2  TPT general  ... v2 = ic2 + g v1;  out = m0 v0 + m1 v1 + m2 v2;  ic2' = 2 v2 - ic2
//   KIND 0  literal processBandStereo form (coefficients that are not TPT-consistent)
// SAT: fused saturation  out*(1-s) + tanh27/9(out)*s  ==  out * (alpha + gamma / (out^2 + 3)),  alpha = (9-8s)/9,
// gamma = 8s/3, with 1/(out^2+3) from the hardware seed (2^-23) and one Newton step (2^-46); the correction term is
// <= 18 % of the result.  hiMax tracks the high word of d = out^2 + 3 (SAT; d >= 23.25 <=> |out| >= 4.5 up to rounding,
// NaN/Inf compare high) or of |out| (no SAT).
// FUSE: pass 1 of the next band (c += w_next[j] * out[j], two accumulator pairs) rides along, so its independent DFMAs and
// weight loads fill the issue slots the recurrence's dependency chain leaves empty.
struct EqAcc { double c1, c2, d1, d2; };
template <bool SAT, int KIND, bool FUSE = false>
__device__ __forceinline__ void eq_pass2(double (&x)[kEqL], double& ic1, double& ic2, const double* __restrict__ bc, double alpha,
                                         double gamma, unsigned& hiMax, const double* __restrict__ wNext = nullptr, EqAcc* acc = nullptr)
{
    const double a1 = bc[0], a2 = bc[1], a3 = bc[2], m0 = bc[3], m1 = bc[4], m2 = bc[5], g = bc[8], g2 = bc[9];
#pragma unroll
    for (int j = 0; j < kEqL; ++j)
    {
        const double v0 = x[j];
        double out;
        if (KIND == 1)
        {
            const double v1 = fma(a1, ic1, fma(-a2, ic2, a2 * v0));
            out = fma(m1, v1, v0);
            ic1 = fma(2.0, v1, -ic1);
            ic2 = fma(g2, v1, ic2);
        }
        else if (KIND == 2)
        {
            const double v1 = fma(a1, ic1, fma(-a2, ic2, a2 * v0));
            const double v2 = fma(g, v1, ic2);
            out = fma(m0, v0, fma(m1, v1, m2 * v2));
            ic1 = fma(2.0, v1, -ic1);
            ic2 = fma(2.0, v2, -ic2);
        }
        else
        {
            const double v3 = v0 - ic2;
            const double v1 = fma(a1, ic1, a2 * v3);
            const double v2 = fma(a2, ic1, fma(a3, v3, ic2));
            ic1 = fma(2.0, v1, -ic1);
            ic2 = fma(2.0, v2, -ic2);
            out = fma(m0, v0, fma(m1, v1, m2 * v2));
        }
        if (SAT)
        {
            const double d = fma(out, out, 3.0);
            hiMax = max(hiMax, (unsigned) __double2hiint(d));
            double r0;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
            const double e = fma(-d, r0, 1.0);
            const double r = fma(r0, e, r0);
            x[j] = out * fma(gamma, r, alpha);
        }
        else
        {
            hiMax = max(hiMax, (unsigned) __double2hiint(out) & 0x7fffffffu);
            x[j] = out;
        }
        if (FUSE)
        {
            const double2 w = reinterpret_cast<const double2*>(wNext)[j];
            if (j & 1)
            {
                acc->d1 = fma(w.x, x[j], acc->d1);
                acc->d2 = fma(w.y, x[j], acc->d2);
            }
            else
            {
                acc->c1 = fma(w.x, x[j], acc->c1);
                acc->c2 = fma(w.y, x[j], acc->c2);
            }
        }
    }
}

// Pass 2 as two interleaved chains: samples [0, L/2) from the block's start state and [L/2, L) from the state half a block
// later (A^(L/2) s + the first half's zero-state response, which pass 1 accumulates separately anyway).  One recurrence is
// a chain of three dependent DFMAs per sample (8 cycles each); two independent chains per thread halve the time a warp
// spends waiting on its own results (the kernel's dominant stall, profiles/r01h_stalls_eq_kernel.txt).
#ifndef CPQ_EQ_UW
#define CPQ_EQ_UW 1
#endif
template <bool SAT, int KIND>
__device__ __forceinline__ void eq_pass2x2(double (&x)[kEqL], double sA1, double sA2, double sB1, double sB2, const double* __restrict__ bc,
                                           double alpha, double gamma, unsigned& hiMax)
{
    const double a1 = bc[0], a2 = bc[1], a3 = bc[2], m0 = bc[3], m1 = bc[4], m2 = bc[5], g = bc[8], g2 = bc[9];
    constexpr int H = kEqL / 2;
    auto sat = [&](double out) -> double {
        if (SAT)
        {
            const double d = fma(out, out, 3.0);
            hiMax = max(hiMax, (unsigned) __double2hiint(d));
            double r0;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
            const double e = fma(-d, r0, 1.0);
            const double r = fma(r0, e, r0);
            return out * fma(gamma, r, alpha);
        }
        hiMax = max(hiMax, (unsigned) __double2hiint(out) & 0x7fffffffu);
        return out;
    };
#if CPQ_EQ_UW
    if (KIND == 1)
    {
        // Peaking band in the rotated state  u = a1 ic1 - a2 ic2,  w = a1 ic1 + a2 ic2:  v1 = u + a2 v0 and, with
        // ic1' = 2 v1 - ic1, ic2' = ic2 + 2 g v1 (a3 = g a2):  u' = 2 (a1 - a3) v1 - w,  w' = 2 (a1 + a3) v1 - u  -- four FP64
        // operations per sample instead of six.  u and w nearly cancel in ic2 for low bands (a2 << a1), which a long run would
        // amplify, but a chain here is half a thread block (16 samples) from a start state the scan supplies in the (ic1, ic2)
        // basis: measured <= 1e-13 of the long-double recurrence at |y| ~ 5 (2.7e-12 at |y| = 86, Q = 0.01, +48 dB), against
        // 1e-14 for the (ic1, ic2) form -- a factor the 1e-10 budget absorbs (tests/test_gpu_parity.py stress cases).
        const double cu = bc[10], cw = bc[11];
        const double tA = a2 * sA2, tB = a2 * sB2;
        double uA = fma(a1, sA1, -tA), wA = fma(a1, sA1, tA), uB = fma(a1, sB1, -tB), wB = fma(a1, sB1, tB);
#pragma unroll
        for (int j = 0; j < H; ++j)
        {
            const double vA = x[j], vB = x[j + H];
            const double v1A = fma(a2, vA, uA), v1B = fma(a2, vB, uB);
            const double nA = fma(cu, v1A, -wA), nB = fma(cu, v1B, -wB);
            wA = fma(cw, v1A, -uA);
            wB = fma(cw, v1B, -uB);
            uA = nA;
            uB = nB;
            x[j] = sat(fma(m1, v1A, vA));
            x[j + H] = sat(fma(m1, v1B, vB));
        }
        return;
    }
#endif
    auto step = [&](double v0, double& ic1, double& ic2) -> double {
        double out;
        if (KIND == 1)
        {
            const double v1 = fma(a1, ic1, fma(-a2, ic2, a2 * v0));
            out = fma(m1, v1, v0);
            ic1 = fma(2.0, v1, -ic1);
            ic2 = fma(g2, v1, ic2);
        }
        else if (KIND == 2)
        {
            const double v1 = fma(a1, ic1, fma(-a2, ic2, a2 * v0));
            const double v2 = fma(g, v1, ic2);
            out = fma(m0, v0, fma(m1, v1, m2 * v2));
            ic1 = fma(2.0, v1, -ic1);
            ic2 = fma(2.0, v2, -ic2);
        }
        else
        {
            const double v3 = v0 - ic2;
            const double v1 = fma(a1, ic1, a2 * v3);
            const double v2 = fma(a2, ic1, fma(a3, v3, ic2));
            ic1 = fma(2.0, v1, -ic1);
            ic2 = fma(2.0, v2, -ic2);
            out = fma(m0, v0, fma(m1, v1, m2 * v2));
        }
        return sat(out);
    };
#pragma unroll
    for (int j = 0; j < H; ++j)
    {
        x[j] = step(x[j], sA1, sA2);
        x[j + H] = step(x[j + H], sB1, sB2);
    }
}

#ifndef CPQ_EQ_SLEEP
#define CPQ_EQ_SLEEP 40
#endif
#ifndef CPQ_EQ_MBAR
#define CPQ_EQ_MBAR 1      // mailbox flags are mbarriers (arrive / try_wait: a waiting warp is suspended by the hardware) instead of polled words
#endif
__device__ __forceinline__ void eq_mbar_init(unsigned long long* bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned) __cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void eq_mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((unsigned) __cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void eq_mbar_wait(unsigned long long* bar)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "EQW_%=:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], 0, 0x989680;\n"
        "@p bra EQD_%=;\n"
        "bra EQW_%=;\n"
        "EQD_%=:\n"
        "}\n" ::"r"((unsigned) __cvta_generic_to_shared(bar)) : "memory");
}
#ifndef CPQ_EQ_ILP2
#define CPQ_EQ_ILP2 1
#endif
#ifndef CPQ_EQ_MINBLOCKS
#define CPQ_EQ_MINBLOCKS (CPQ_EQ_L == 16 ? 3 : 2)
#endif
// POST = the launch runs output stages / the output clamp; the plain conv -> EQ -> gain launch carries none of that code.
// PAR  = FilterStructure::Parallel: every band filters the tile *input* and the differences are summed (a second block of
//        registers per thread, hence one CTA per SM).   STATS = the launch accumulates the AGC block statistics.
// LB   = look-back links instead of chained links (see bandStart): for launches with few sequences, where the links between
//        the tiles of one sequence are the critical path.  A separate instantiation so that neither form carries the other's code.
template <bool POST, bool PAR = false, bool STATS = false, bool LB = false>
__global__ void __launch_bounds__(kEqThreads, PAR ? 1 : CPQ_EQ_MINBLOCKS) eq_kernel(EqArgs a)
{
    const unsigned postMask = POST ? a.postMask : 0u;
    // band constants in shared memory: the chained form has no use for the Pw table at the end of each band's block
    constexpr int kStr = LB ? kEqcStride : kEqcPw;
    extern __shared__ __align__(16) double eq_smem[];
    double* tile = eq_smem;                                        // [7 warps][32 lanes][18]
    double* cst = tile + kEqCThreads * kEqPad;                     // this sequence's band constants
    double2* agg = reinterpret_cast<double2*>(cst + CPQ_NUM_BANDS * kStr);  // [stage][8] zero-state response of warp w's segment
    double2* tin = agg + kEqStages * 8;                                           // [stage] state at the start of the tile
    unsigned long long* mbA = reinterpret_cast<unsigned long long*>(tin + kEqStages);   // [stage][8] "aggregate posted" (one phase each)
    unsigned long long* mbT = mbA + kEqStages * 8;                                // [stage] "tile-in state posted"
    unsigned* sTicket = reinterpret_cast<unsigned*>(mbT + kEqStages);
    double2* endSt = reinterpret_cast<double2*>(sTicket + 4);                     // look-back form: [stage][8] state after a serially run segment
    unsigned long long* mbE = reinterpret_cast<unsigned long long*>(endSt + kEqMailLb);   // ... and "end state posted"
    double* cstPost = eq_smem + eqSmemDoubles(LB);                                   // output-stage constants (present iff postMask)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) *sTicket = atomicAdd(a.chain.ticket, 1u);
    for (int i = tid; i < kEqMail; i += kEqThreads) eq_mbar_init(mbA + i);   // mbA and mbT are contiguous
    if (LB)
        for (int i = tid; i < kEqMailLb; i += kEqThreads) eq_mbar_init(mbE + i);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const unsigned ticket = *sTicket;
    // run-major ticket order: the predecessor (same sequence, previous tile) always holds a smaller ticket
    const int run = (int) (ticket / (unsigned) a.nSeq);
    const int seq = (int) (ticket % (unsigned) a.nSeq);
    if (run >= a.nRuns) return;

    double* io = a.io + (size_t) seq * a.ioStride;
    const int set = (a.doEq || a.doGain) ? a.setOfSeq[seq] : 0;
    // stages this sequence runs: its active EQ bands (bits 0..19) and the enabled output stages (bits 20..23)
    const unsigned bandBits = a.doEq ? a.bandMask[seq] : 0u;   // bits 0..19 bands, bit 31 = Parallel structure
    const unsigned mask = (bandBits & ((bandBits >> 31) ? a.bandSelectPar : a.bandSelect) & ((1u << CPQ_NUM_BANDS) - 1u)) | (postMask << CPQ_NUM_BANDS);
    const unsigned scalarBits = a.doEq ? ((a.scalarMask ? a.scalarMask[seq] : 0u) | a.scalarAll) : 0u;
    const int64_t t0 = (int64_t) run * kEqTile;

    if (a.doEq)
    {
        const double* __restrict__ src = a.eqc + (size_t) set * CPQ_NUM_BANDS * kEqcStride;
        for (int i = tid; i < CPQ_NUM_BANDS * kStr / 2; i += kEqThreads)
        {
            const int b = i / (kStr / 2), o = i - b * (kStr / 2);
            reinterpret_cast<double2*>(cst)[i] = __ldg(reinterpret_cast<const double2*>(src + (size_t) b * kEqcStride) + o);
        }
    }
    if (postMask)
        for (int i = tid; i < kEqPostStages * kStr / 2; i += kEqThreads)
        {
            const int b = i / (kStr / 2), o = i - b * (kStr / 2);
            reinterpret_cast<double2*>(cstPost)[i] = __ldg(reinterpret_cast<const double2*>(a.postc + (size_t) b * kEqcStride) + o);
        }
    __syncthreads();   // constants + cleared flags visible; the only CTA-wide barrier besides the ticket

    // ---- tile-to-tile chain: records (seq, tile, band) in global memory, written by the last warp of a tile and read by
    // the first warp of the next one.  The read of band b+1's record is issued while band b is computed, so its L2 round
    // trip is never on the critical path; the predecessor tile started earlier (ticket order) and is normally far ahead.
    const bool lastTile = run + 1 >= a.nRuns;
    const bool fullLast = lastTile && (a.T - t0) == kEqTile;
    const double2* recIn = (mask && run > 0) ? a.chain.rec + (size_t) ((size_t) seq * a.nRuns + (run - 1)) * kEqStages : nullptr;
    double2* recOut = (mask && !lastTile) ? a.chain.rec + (size_t) ((size_t) seq * a.nRuns + run) * kEqStages : nullptr;
    auto nextBand = [&](int b) { ++b; while (b < kEqStages && !((mask >> b) & 1u)) ++b; return b; };
    // final state of stage b (EQ bands -> stateOut, output stages -> postStateOut)
    auto storeFinal = [&](int b, double s1, double s2) {
        double* dst = b < CPQ_NUM_BANDS ? (a.stateOut ? a.stateOut + ((size_t) seq * CPQ_NUM_BANDS + b) * 2 : nullptr)
                                        : (a.postStateOut ? a.postStateOut + ((size_t) seq * kEqPostStages + (b - CPQ_NUM_BANDS)) * 2 : nullptr);
        if (dst) { dst[0] = s1; dst[1] = s2; }
    };
    double2 pre = make_double2(0.0, 0.0);   // warp 0: prefetched inbound record of band preBand
    int preBand = -1;
    if (warp == 0 && recIn)
    {
        preBand = nextBand(-1);
        if (preBand < kEqStages) pre = ld_volatile_f64x2(recIn + preBand);
    }

    // ================= compute warps: one 512-sample segment each =================
    const double sat = a.doEq ? a.sat[set] : 0.0;
    const int bmask = a.blockLog2 >= 0 ? (1 << a.blockLog2) - 1 : 0;   // fast path only (power-of-two blocks)
    const int64_t w0 = t0 + (int64_t) warp * kEqSeg;                       // first sample of this warp's segment
    const int nValid = (int) max((int64_t) 0, min((int64_t) kEqSeg, a.T - w0));
    double* wtile = tile + warp * 32 * kEqPad;
    double* myStash = wtile + lane * kEqPad;   // 16 consecutive slots + 2 pad slots

    // ---- coalesced load + layer assembly (Get) ----
    // Fast path: callbacks are whole multiples of 512 samples and the tail streams are stored in stream order
    // (regular plans), so the tail source position is one lookup per (layer, 512-sample sub-block).
    const bool segFast = a.assemble && a.blockLog2 >= 9 && a.blockMap[0] == nullptr && a.blockMap[1] == nullptr && a.l0 == nullptr;
    auto cbOf = [&](int64_t t) -> int64_t { return a.blockLog2 >= 0 ? (t >> a.blockLog2) : t / a.blockSize; };
    if (!a.assemble || segFast)
    {
        const bool outer = a.assemble && a.outer;
        const double wet = a.wetGain;
#pragma unroll
        for (int sb = 0; sb < kEqSeg / 512; ++sb)
        {
            const int64_t ws = w0 + sb * 512;
            const double* tp[2] = { nullptr, nullptr };
            double tg[2] = { 0.0, 0.0 };
            if (segFast && sb * 512 < nValid)
            {
                const int64_t c = ws >> a.blockLog2;
                const int64_t off = ws & (int64_t) bmask;
#pragma unroll
                for (int l = 0; l < 2; ++l)
                    if (l < a.nTail)
                    {
                        const int64_t sp = __ldg(a.tailSrc[l] + c);
                        if (sp >= 0)
                        {
                            tp[l] = a.tail[l] + (size_t) seq * a.tailStride[l] + sp + off;
                            tg[l] = a.tailGain[l];
                        }
                    }
            }
            // Loads are issued in unconditional batches of eight per array (clamped index instead of a branch), so one
            // memory round trip serves eight samples; a branch per sample would serialise them.
            const double* ip = io + ws;
            const int nv = nValid - sb * 512;            // valid samples of this sub-block (uniform per warp)
            if (nv <= 0)
            {
#pragma unroll
                for (int k = 0; k < 16; ++k) wtile[eq_sidx(sb * 512 + lane + 32 * k)] = 0.0;
                continue;
            }
            // 16-byte path: the whole 512-sample sub-block of every array in one batch of eight LDG.128 per lane
            const bool vec16 = (nv & 1) == 0 && ((reinterpret_cast<uintptr_t>(ip) | reinterpret_cast<uintptr_t>(tp[0]) | reinterpret_cast<uintptr_t>(tp[1])) & 15) == 0;
            if (vec16)
            {
                const double2* ip2 = reinterpret_cast<const double2*>(ip);
                const double2* t02 = reinterpret_cast<const double2*>(tp[0]);
                const int64_t ioOff = ip - a.io;   // element offset of this sub-block, the same in every peer's buffer
                const double2* t12 = reinterpret_cast<const double2*>(tp[1]);
                const int last2 = nv / 2 - 1;
                double2 va[8], vb[8], vc[8];
                if (STATS && a.nPeers > 0)
                {
#pragma unroll
                    for (int k = 0; k < 8; ++k) va[k] = make_double2(0.0, 0.0);
                    for (int p = 0; p < a.nPeers; ++p)
                    {
                        const double2* pp = reinterpret_cast<const double2*>(a.peer[p] + ioOff);
                        double2 vp[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) vp[k] = pp[min(lane + 32 * k, last2)];
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                        {
                            va[k].x = p == 0 ? vp[k].x : va[k].x + vp[k].x;
                            va[k].y = p == 0 ? vp[k].y : va[k].y + vp[k].y;
                        }
                    }
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < 8; ++k) va[k] = ip2[min(lane + 32 * k, last2)];
                }
                if (tp[0])
                {
#pragma unroll
                    for (int k = 0; k < 8; ++k) vb[k] = __ldg(t02 + min(lane + 32 * k, last2));
                }
                if (tp[1])
                {
#pragma unroll
                    for (int k = 0; k < 8; ++k) vc[k] = __ldg(t12 + min(lane + 32 * k, last2));
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                {
                    const int i2 = lane + 32 * k;
                    double2 v = va[k];
                    if (tp[0]) { v.x += vb[k].x * tg[0]; v.y += vb[k].y * tg[0]; }
                    if (tp[1]) { v.x += vc[k].x * tg[1]; v.y += vc[k].y * tg[1]; }
                    if (outer)
                    {
                        if (!(fabs(v.x) < 1.0e300)) v.x = 0.0;
                        if (!(fabs(v.y) < 1.0e300)) v.y = 0.0;
                        v.x *= wet;
                        v.y *= wet;
                    }
                    if (i2 > last2) v = make_double2(0.0, 0.0);
                    *reinterpret_cast<double2*>(wtile + eq_sidx(sb * 512 + 2 * i2)) = v;
                }
                continue;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                double va[8], vb[8], vc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                {
                    const int idx = min(lane + 32 * (8 * h + k), nv - 1);
                    va[k] = ip[idx];
                    if (STATS && a.nPeers > 0)
                    {
                        va[k] = a.peer[0][(ip - a.io) + idx];
                        for (int p = 1; p < a.nPeers; ++p) va[k] += a.peer[p][(ip - a.io) + idx];
                    }
                }
                if (tp[0])
                {
#pragma unroll
                    for (int k = 0; k < 8; ++k) vb[k] = __ldg(tp[0] + min(lane + 32 * (8 * h + k), nv - 1));
                }
                if (tp[1])
                {
#pragma unroll
                    for (int k = 0; k < 8; ++k) vc[k] = __ldg(tp[1] + min(lane + 32 * (8 * h + k), nv - 1));
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                {
                    const int i = lane + 32 * (8 * h + k);
                    double v = va[k];
                    if (tp[0]) v += vb[k] * tg[0];
                    if (tp[1]) v += vc[k] * tg[1];
                    if (outer)
                    {
                        if (!(fabs(v) < 1.0e300)) v = 0.0;
                        v *= wet;
                    }
                    if (i >= nv) v = 0.0;
                    wtile[eq_sidx(sb * 512 + i)] = v;
                }
            }
        }
    }
    else
    {
        // generic path: per-sample callback lookup (block < 512) and block-mapped tail streams (irregular plans)
        for (int i = lane; i < kEqSeg; i += 32)
        {
            double v = 0.0;
            if (i < nValid)
            {
                const int64_t t = w0 + i;
                v = io[t];
                if (STATS && a.nPeers > 0)
                {
                    v = a.peer[0][(io - a.io) + t];
                    for (int p = 1; p < a.nPeers; ++p) v += a.peer[p][(io - a.io) + t];
                }
                const int64_t c = cbOf(t);
                const int off = (int) (t - c * a.blockSize);
                // Beyond what the L0 ring delivered the caller zero-fills the WHOLE output of the callback, tails included
                // (StereoConvolver::process: got = Get(out, n); memset(out + got, 0, n - got), Runtime.cpp:1174-1183)
                const bool live = !a.l0 || off < __ldg(a.l0Count + c);
                if (a.l0) v = live ? __ldg(a.l0 + (size_t) seq * a.l0Stride + __ldg(a.l0Src + c) + off) : 0.0;
                for (int l = 0; l < (live ? a.nTail : 0); ++l)
                {
                    const int64_t sp = __ldg(a.tailSrc[l] + c);
                    if (sp >= 0)
                    {
                        int64_t pos = sp + off;
                        if (a.blockMap[l])
                        {
                            const int64_t j = pos >> a.tailPartLog2[l];
                            pos = ((int64_t) __ldg(a.blockMap[l] + j) << a.tailPartLog2[l]) + (pos - (j << a.tailPartLog2[l]));
                        }
                        const double tv = __ldg(a.tail[l] + (size_t) seq * a.tailStride[l] + pos);
                        v += tv * a.tailGain[l];
                    }
                }
                if (a.outer)
                {
                    if (!(fabs(v) < 1.0e300)) v = 0.0;
                    v *= a.wetGain;
                }
            }
            wtile[eq_sidx(i)] = v;
        }
    }
    __syncwarp();

    double x[kEqL];
    auto loadBlock = [&](unsigned& hi) {
#pragma unroll
        for (int j = 0; j < kEqL / 2; ++j)
        {
            const double2 v = reinterpret_cast<const double2*>(myStash)[j];
            x[2 * j] = v.x;
            x[2 * j + 1] = v.y;
            hi = max(hi, max((unsigned) __double2hiint(v.x) & 0x7fffffffu, (unsigned) __double2hiint(v.y) & 0x7fffffffu));
        }
    };
    unsigned hiIn = 0;   // max of |x|'s high word over the raw input
    loadBlock(hiIn);
    // AGC block statistics: sum of squares per callback.  A thread's block never straddles a callback (block >= 64); the
    // lanes of one callback are reduced by shuffles and one lane adds the partial sum (several warps share a callback
    // only when block > 1024).
    auto blockStats = [&](double* dst) {
        double sq = 0.0;
#pragma unroll
        for (int j = 0; j < kEqL; ++j) sq = fma(x[j], x[j], sq);
        const int64_t tb = w0 + (int64_t) lane * kEqL;
        if (a.blockLog2 < 0)
        {
            // host block not a power of two: the lanes of a callback are not an aligned group, and a thread's block may
            // straddle two callbacks
            if (tb >= a.T) return;
            const int64_t cb0 = tb / a.blockSize;
            const int split = (int) ((cb0 + 1) * a.blockSize - tb);
            if (split >= kEqL) atomicAdd(dst + (size_t) seq * a.nCallbacks + cb0, sq);
            else
            {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int j = 0; j < kEqL; ++j)
                {
                    if (j < split) s0 = fma(x[j], x[j], s0);
                    else s1 = fma(x[j], x[j], s1);
                }
                atomicAdd(dst + (size_t) seq * a.nCallbacks + cb0, s0);
                if (cb0 + 1 < a.nCallbacks) atomicAdd(dst + (size_t) seq * a.nCallbacks + cb0 + 1, s1);
            }
            return;
        }
        const int lanesPerCb = min(32, (1 << a.blockLog2) / kEqL);
        for (int o = 1; o < lanesPerCb; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if ((lane & (lanesPerCb - 1)) == 0 && tb < a.T) atomicAdd(dst + (size_t) seq * a.nCallbacks + (tb >> a.blockLog2), sq);
    };
    if (STATS && a.sumsqIn) blockStats(a.sumsqIn);

    // gains the reference applies between the EQ bands and the output stages are applied in registers when such a stage
    // runs: total gain (Processing.cpp:1262-1274) before OutputFilter / DC blocker, makeup (DSPCoreDouble.cpp:465-469)
    // before the DC blocker; the store stage then skips them
    const bool gainInReg = a.doGain && (postMask != 0u);
    const int gainRow = a.gainBySeq > 0 ? seq / a.gainBySeq : set;
    const bool makeupInReg = a.doEpilogue && ((postMask >> 3) & 1u);
    if (mask)
    {
        const double alpha = fma(-8.0, sat, 9.0) / 9.0, gamma = 8.0 * sat / 3.0;
        const unsigned thrHi = sat > 0.0 ? 0x40374000u : 0x40590000u;   // high words of 4.5^2 + 3 / of 100.0
        // the thread whose block starts at sample T holds the sequence's final state when the last tile is partial
        const int64_t remT = a.T - t0;
        const bool ownsFinal = remT < kEqTile && remT >= 0 && tid == (int) (remT / kEqL);
        const int finalOff = ownsFinal ? (int) (remT % kEqL) : 0;   // T need not be a multiple of the thread block (441-sample hosts)
        // The owner's final state: the state at the start of its block advanced over the finalOff samples the sequence still has
        // in it (x holds the stage's *input* here; the band state does not depend on the saturated output).
        auto storeOwned = [&](int b, const double* __restrict__ bc, double s1, double s2) {
            if (finalOff > 0)
            {
                const int kind = (int) bc[7];
#pragma unroll
                for (int j = 0; j < kEqL - 1; ++j)
                    if (j < finalOff)
                    {
                        const double v0 = x[j];
                        if (kind == 3)        // DF2T biquad (postPass2)
                        {
                            const double y = fma(bc[0], v0, s1);
                            s1 = fma(bc[1], v0, fma(-bc[3], y, s2));
                            s2 = fma(-bc[4], y, bc[2] * v0);
                        }
                        else if (kind == 4)   // two one-pole sections
                        {
                            s1 = fma(bc[0], v0 - s1, s1);
                            const double v = v0 - s1;
                            s2 = fma(bc[1], v - s2, s2);
                        }
                        else                  // TPT SVF, literal form
                        {
                            const double v3 = v0 - s2;
                            const double v1 = fma(bc[0], s1, bc[1] * v3);
                            const double v2 = fma(bc[1], s1, fma(bc[2], v3, s2));
                            s1 = fma(2.0, v1, -s1);
                            s2 = fma(2.0, v2, -s2);
                        }
                    }
            }
            storeFinal(b, s1, s2);
        };

        // Everything of one band up to the start state of this thread's block.  `link`: take part in the chain (wait for
        // the mailbox, post the successor's); a replayed band finds its mailbox already filled and posts nothing.
        double mid1 = 0.0, mid2 = 0.0;   // state half a block after (ic1, ic2): start of the second pass-2 chain
        // One band over this warp's segment the way the reference runs it, sample after sample: lane l continues from the
        // state lane l - 1 ended with (state resets included, eq_pass2_exact).  In: (s1, s2) the state at the start of the
        // segment; out: the state after it, (b1, b2) the state at the start of this thread's block, x the band output.
        auto serialBand = [&](const double* __restrict__ bc, bool scalarTanh, double& s1, double& s2, double& b1, double& b2) {
#pragma unroll 1
            for (int l = 0; l < 32; ++l)
            {
                if (lane == l)
                {
                    b1 = s1;
                    b2 = s2;
                    if (ownsFinal && finalOff > 0) eq_pass2_exact(x, s1, s2, bc, sat, scalarTanh, finalOff, &b1, &b2);   // the owner keeps the state at sample T
                    else eq_pass2_exact(x, s1, s2, bc, sat, scalarTanh);
                }
                s1 = __shfl_sync(0xffffffffu, s1, l);
                s2 = __shfl_sync(0xffffffffu, s2, l);
            }
        };
        // `risk`: the band's input or coefficients may drive a state out of range within this segment (warp-uniform; the same
        // value when a band is replayed).  What happens then is the call site's `how` (a compile-time tag, so that the hot
        // loop's instantiation carries none of the serial code): kEqNoSerial = never (output stages), kEqBailOut = return true
        // before anything is posted (fast mode: the caller switches to exact mode and comes back), kEqSerial = run the band
        // serially, post the true state after the segment and return true: x then already holds the band's output.
        auto bandStart = [&](auto how, int b, const double* __restrict__ bc, bool link, double& ic1, double& ic2, bool risk) -> bool {
            constexpr int kHow = decltype(how)::value;
            bool serial = false;
            if (kHow == kEqBailOut && risk) return true;
            // ---- pass 1: zero-state response of this thread's samples (four accumulation chains) ----
#if CPQ_EQ_ILP2
            // per half block, with the same 16 weights A^(L/2-1-j) b: c = A^(L/2) lo + hi
            double lo1 = 0.0, lo2 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
            for (int j = 0; j < kEqL / 2; ++j)
            {
                const double2 w = reinterpret_cast<const double2*>(bc + kEqcW)[j + kEqL / 2];
                lo1 = fma(w.x, x[j], lo1);
                lo2 = fma(w.y, x[j], lo2);
                c1 = fma(w.x, x[j + kEqL / 2], c1);
                c2 = fma(w.y, x[j + kEqL / 2], c2);
            }
            const double2 mh0 = *reinterpret_cast<const double2*>(bc + kEqcMh), mh1 = *reinterpret_cast<const double2*>(bc + kEqcMh + 2);
            c1 = fma(mh0.x, lo1, fma(mh0.y, lo2, c1));
            c2 = fma(mh1.x, lo1, fma(mh1.y, lo2, c2));
#else
            double c1 = 0.0, c2 = 0.0, d1 = 0.0, d2 = 0.0;
#pragma unroll
            for (int j = 0; j < kEqL; j += 2)
            {
                const double2 wa = reinterpret_cast<const double2*>(bc + kEqcW)[j];
                const double2 wb = reinterpret_cast<const double2*>(bc + kEqcW)[j + 1];
                c1 = fma(wa.x, x[j], c1);
                c2 = fma(wa.y, x[j], c2);
                d1 = fma(wb.x, x[j + 1], d1);
                d2 = fma(wb.y, x[j + 1], d2);
            }
            c1 += d1;
            c2 += d2;
#endif
            // ---- warp inclusive scan of s -> A^L s + c ----
#pragma unroll
            for (int d = 0; d < 5; ++d)
            {
                const double p1 = __shfl_up_sync(0xffffffffu, c1, 1 << d);
                const double p2 = __shfl_up_sync(0xffffffffu, c2, 1 << d);
                if (lane >= (1 << d))
                {
                    const double2 r0 = reinterpret_cast<const double2*>(bc + kEqcMs + 4 * d)[0];
                    const double2 r1 = reinterpret_cast<const double2*>(bc + kEqcMs + 4 * d)[1];
                    c1 = fma(r0.x, p1, fma(r0.y, p2, c1));
                    c2 = fma(r1.x, p1, fma(r1.y, p2, c2));
                }
            }
            // exclusive value (state contribution before this thread, relative to the segment start)
            double e1 = __shfl_up_sync(0xffffffffu, c1, 1);
            double e2 = __shfl_up_sync(0xffffffffu, c2, 1);
            if (lane == 0) { e1 = 0.0; e2 = 0.0; }

            // ---- link by look-back: every warp posts the zero-state response of its segment as soon as its scan is done
            // (it depends on nobody), warp 0 posts the state at the start of the tile (from the previous tile's record), and
            // warp w composes  s_w = A^(1024 w) s_tile + sum_{i<w} A^(1024 (w-1-i)) agg_i  itself.  No warp waits for another
            // warp's *state*, only for aggregates that all warps produce at about the same time, so the warps of a tile run
            // side by side instead of one link latency behind each other, and the state that leaves the tile is ready one
            // matrix-vector product after the state that enters it (a serial chain of eight links before). ----
            // state at the start of the tile (warp 0 only): Reset, the carried state of a streaming call, or the previous tile's record
            auto tileIn = [&]() -> double2 {
                double2 sv = make_double2(0.0, 0.0);   // Reset state for the first tile
                if (!recIn && run == 0)
                {
                    // streaming continuation: the state the previous call left behind
                    const double* si = b < CPQ_NUM_BANDS ? (a.stateIn ? a.stateIn + ((size_t) seq * CPQ_NUM_BANDS + b) * 2 : nullptr)
                                                         : (a.postStateIn ? a.postStateIn + ((size_t) seq * kEqPostStages + (b - CPQ_NUM_BANDS)) * 2 : nullptr);
                    if (si) sv = make_double2(ld_cg_f64(si), ld_cg_f64(si + 1));   // the same buffer receives the final state: no read-only path
                }
                if (recIn)
                {
                    sv = (preBand == b) ? pre : ld_volatile_f64x2(recIn + b);
                    while (rec_pending(sv))
                    {
                        __nanosleep(100);
                        sv = ld_volatile_f64x2(recIn + b);
                    }
                    preBand = nextBand(b);
                    if (preBand < kEqStages) pre = ld_volatile_f64x2(recIn + preBand);
                }
                return sv;
            };
            double p1 = 0.0, p2 = 0.0;
            if (LB)
            {
                // A segment that runs serially has no zero-state response to offer: it posts a marker, and later the state after
                // the segment (endSt), from which its successors restart their composition.
                serial = kHow == kEqSerial && risk;
                const double marker = __longlong_as_double(0x7ff8c0de5e71a100ll);
                if (link)
                {
                    if (lane == 31)
                    {
                        agg[b * 8 + warp] = serial ? make_double2(marker, marker) : make_double2(c1, c2);
                        eq_mbar_arrive(mbA + b * 8 + warp);   // release
                    }
                    if (warp == 0)
                    {
                        const double2 sv = tileIn();
                        if (lane == 0)
                        {
                            tin[b] = sv;
                            eq_mbar_arrive(mbT + b);
                        }
                    }
                }
                bool fromTile = true;
                for (int i = 0; i < warp; ++i)
                {
                    if (link) eq_mbar_wait(mbA + b * 8 + i);   // acquire; a replayed band finds every mailbox already filled
                    const double2 ai = agg[b * 8 + i];
                    if (__double_as_longlong(ai.x) == __double_as_longlong(marker))
                    {
                        if (link) eq_mbar_wait(mbE + b * 8 + i);
                        const double2 ei = endSt[b * 8 + i];   // the state after segment i: nothing before it matters any more
                        p1 = ei.x;
                        p2 = ei.y;
                        fromTile = false;
                    }
                    else
                        matvec2(bc + kEqcMw, p1, p2, ai.x, ai.y);  // Horner: zero-state response of segments 0..i at the start of segment i + 1
                }
                if (fromTile)
                {
                    if (link) eq_mbar_wait(mbT + b);
                    const double2 tv = tin[b];
                    double q1 = tv.x, q2 = tv.y;
                    matvec2(bc + kEqcPw + 4 * warp, q1, q2, p1, p2);   // A^(1024 w) s_tile + ...
                    p1 = q1;
                    p2 = q2;
                }
                double o1 = p1, o2 = p2;
                if constexpr (kHow == kEqSerial)
                    if (serial)
                    {
                        serialBand(bc, (scalarBits >> b) & 1u, o1, o2, ic1, ic2);
                        if (link && lane == 31 && warp + 1 < kEqCWarps)
                        {
                            endSt[b * 8 + warp] = make_double2(o1, o2);
                            eq_mbar_arrive(mbE + b * 8 + warp);   // release
                        }
                    }
                if (link && lane == 31 && warp + 1 == kEqCWarps)
                {
                    if (!serial) matvec2(bc + kEqcMw, o1, o2, c1, c2);   // state after the tile
                    if (recOut)
                        st_volatile_f64x2(recOut + b, make_double2(rec_clean(o1), rec_clean(o2)));
                    else if (fullLast)
                        storeFinal(b, o1, o2);
                }
            }
            else
            {
                // ---- link by chain (batches with at least one sequence per SM): warp w takes the state at the start of its
                // segment from mailbox agg[b][w] (here: a state, not an aggregate), posted by lane 31 of warp w - 1 right after
                // its scan.  The warps of a tile run skewed by one link latency, which also keeps them in different phases of
                // the band loop -- measured 14 % faster than the look-back form when the machine is full of independent tiles,
                // and twice as slow on a single stereo stream (BASELINE config 2), where the serial links are all there is. ----
                if (link)
                {
                    if (warp == 0)
                    {
                        const double2 sv = tileIn();
                        if (lane == 0) agg[b * 8] = sv;   // kept for an exact-mode replay of this band
                        __syncwarp();
                    }
                    else
                        eq_mbar_wait(mbA + b * 8 + warp);   // acquire: the poster's store is visible
                }
                {
                    const double2 sv = agg[b * 8 + warp];
                    p1 = sv.x;
                    p2 = sv.y;
                }
                // the state at the start of the segment is known before anything is posted: a start state that could reach the
                // reset threshold within the segment (transient growth of a TPT band is far below 1e6) also runs serially
                if (kHow != kEqNoSerial)   // integer compare on the high words (1e9 = 0x41cdcd65...; NaN / Inf compare high): keeps the FP64 pipe out of the link
                    serial = risk || max((unsigned) __double2hiint(p1) & 0x7fffffffu, (unsigned) __double2hiint(p2) & 0x7fffffffu) >= 0x41cdcd65u;
                if (kHow == kEqBailOut && serial) return true;
                double o1 = p1, o2 = p2;
                if constexpr (kHow == kEqSerial)
                {
                    if (serial) serialBand(bc, (scalarBits >> b) & 1u, o1, o2, ic1, ic2);
                }
                if (!serial && lane == 31) matvec2(bc + kEqcMw, o1, o2, c1, c2);   // state after this segment
                if (link && lane == 31)
                {
                    if (warp + 1 < kEqCWarps)
                    {
                        agg[b * 8 + warp + 1] = make_double2(o1, o2);
                        eq_mbar_arrive(mbA + b * 8 + warp + 1);   // release
                    }
                    else if (recOut)
                        st_volatile_f64x2(recOut + b, make_double2(rec_clean(o1), rec_clean(o2)));   // state after the tile
                    else if (fullLast)
                        storeFinal(b, o1, o2);
                }
            }
            if (kHow == kEqSerial && serial)
            {
                if (ownsFinal) storeFinal(b, ic1, ic2);   // (serialBand left the state at sample T in the owner's ic)
                return true;
            }
            // ---- state at the start of this thread's block: A^(L lane) s_in + e ----
            matvec2(bc + kEqcPlo + 4 * (lane & 7), p1, p2, 0.0, 0.0);    // A^(L (lane & 7)) (state at the start of this thread's block)
            matvec2(bc + kEqcPhi + 4 * (lane >> 3), p1, p2, e1, e2);     // A^(8L (lane >> 3)) ... + e
            ic1 = p1;
            ic2 = p2;
#if CPQ_EQ_ILP2
            mid1 = fma(mh0.x, p1, fma(mh0.y, p2, lo1));
            mid2 = fma(mh1.x, p1, fma(mh1.y, p2, lo2));
#endif
            if (ownsFinal) storeOwned(b, bc, ic1, ic2);
            return false;
        };

        auto preStage = [&](int b, bool& gainDone, bool& makeupDone) {
            if (b >= CPQ_NUM_BANDS && gainInReg && !gainDone)
            {
                gainDone = true;
                if (a.gainTab)
                {
                    // a thread's block lies in one callback when the host block is a multiple of the thread block, and in at
                    // most two otherwise (block >= 64 > kEqL)
                    const int64_t tb = w0 + (int64_t) lane * kEqL;
                    const int64_t cb0 = cbOf(tb);
                    const double2* gt = reinterpret_cast<const double2*>(a.gainTab) + (size_t) gainRow * a.nCallbacks;
                    const double2 g = __ldg(gt + cb0);
                    const int off0 = (int) (tb - cb0 * a.blockSize);
                    const int split = a.blockSize - off0;   // samples of this block that belong to callback cb0
                    const double2 g1 = (split < kEqL && cb0 + 1 < a.nCallbacks) ? __ldg(gt + cb0 + 1) : g;
#pragma unroll
                    for (int j = 0; j < kEqL; ++j)
                        x[j] *= j < split ? fma((double) (off0 + j), g.y, g.x) : fma((double) (j - split), g1.y, g1.x);
                }
                else
                {
                    const double gc = __ldg(a.gainConst + set);
#pragma unroll
                    for (int j = 0; j < kEqL; ++j) x[j] *= gc;
                }
            }
            if (b == kEqStageDc && makeupInReg && !makeupDone)
            {
                makeupDone = true;
#pragma unroll
                for (int j = 0; j < kEqL; ++j) x[j] *= a.makeup;
            }
        };
        // pass 2 of the linear output stages (no clamps or scrubs in the reference: one code path for fast and exact mode)
        auto postPass2 = [&](const double* __restrict__ bc, double& s1, double& s2) {
            if ((int) bc[7] == 3)
            {
                // DF2T biquad with the reference's FMA association (OutputFilter.cpp:118-137)
                const double b0 = bc[0], b1 = bc[1], b2 = bc[2], a1 = bc[3], a2 = bc[4];
#pragma unroll
                for (int j = 0; j < kEqL; ++j)
                {
                    const double xi = x[j];
                    const double y = fma(b0, xi, s1);
                    s1 = fma(b1, xi, fma(-a1, y, s2));
                    s2 = fma(-a2, y, b2 * xi);
                    x[j] = y;
                }
            }
            else
            {
                // two one-pole sections (UltraHighRateDCBlocker.h:98-126)
                const double al0 = bc[0], al1 = bc[1];
#pragma unroll
                for (int j = 0; j < kEqL; ++j)
                {
                    double v = x[j];
                    s1 = fma(al0, v - s1, s1);
                    v -= s1;
                    s2 = fma(al1, v - s2, s2);
                    x[j] = v - s2;
                }
            }
        };

        // ---- fast mode: bands in order until some lane leaves the regime where the reference's clamps are identities ----
        bool exactMode = a.doEq && __any_sync(0xffffffffu, hiIn >= 0x41cdcd65u);   // |x| >= 1e9 (or NaN/Inf) in the raw input
        const bool rawBig = exactMode;   // ... which may reset the state of the band that filters it (-> serialBand)
        int linked = -1;                                                            // last band whose link this warp has served
        if (PAR && (bandBits >> 31))
        {
            // Parallel structure: every band starts from the tile input (kept in shared memory), so the fast / exact decision
            // is per band; acc follows the reference's two roundings per band, accum = (accum + work) - src (:1194-1197).
            double acc[kEqL];
#pragma unroll
            for (int j = 0; j < kEqL; ++j) acc[j] = 0.0;
            bool first = true;
            for (int b = 0; b < CPQ_NUM_BANDS; ++b)
            {
                if (!((mask >> b) & 1u)) continue;
                const double* __restrict__ bc = cst + b * kStr;
                unsigned dummy = 0;
                if (!first) loadBlock(dummy);
                first = false;
                double ic1, ic2;
                const bool ranSerial = bandStart(std::integral_constant<int, kEqSerial> {}, b, bc, true, ic1, ic2, rawBig || bc[6] != 0.0);   // every band filters the raw input
                const double s1 = ic1, s2 = ic2;
                unsigned hiMax = max((unsigned) __double2hiint(ic1) & 0x7fffffffu, (unsigned) __double2hiint(ic2) & 0x7fffffffu);
                bool rare = exactMode | (hiMax >= 0x426d1a94u) | (bc[6] != 0.0);
                if (ranSerial) rare = false;
                else if (!__any_sync(0xffffffffu, rare))
                {
                    hiMax = 0;
                    const int kind = (int) bc[7];
#if CPQ_EQ_ILP2
                    if (sat > 0.0)
                    {
                        if (kind == 1) eq_pass2x2<true, 1>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                        else if (kind == 2) eq_pass2x2<true, 2>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                        else eq_pass2x2<true, 0>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                    }
                    else
                    {
                        if (kind == 1) eq_pass2x2<false, 1>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                        else if (kind == 2) eq_pass2x2<false, 2>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                        else eq_pass2x2<false, 0>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                    }
#else
                    if (sat > 0.0)
                    {
                        if (kind == 1) eq_pass2<true, 1>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                        else if (kind == 2) eq_pass2<true, 2>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                        else eq_pass2<true, 0>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                    }
                    else
                    {
                        if (kind == 1) eq_pass2<false, 1>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                        else if (kind == 2) eq_pass2<false, 2>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                        else eq_pass2<false, 0>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                    }
#endif
                    rare = hiMax >= thrHi;
                }
                if (!ranSerial && __any_sync(0xffffffffu, rare))
                {
                    loadBlock(dummy);
                    ic1 = s1;
                    ic2 = s2;
                    if (eq_pass2_exact(x, ic1, ic2, bc, sat, (scalarBits >> b) & 1u)) atomicExch(a.fault, 1u);
                }
#pragma unroll
                for (int j = 0; j < kEqL / 2; ++j)
                {
                    const double2 v = reinterpret_cast<const double2*>(myStash)[j];
                    acc[2 * j] = __dadd_rn(__dadd_rn(acc[2 * j], x[2 * j]), -v.x);
                    acc[2 * j + 1] = __dadd_rn(__dadd_rn(acc[2 * j + 1], x[2 * j + 1]), -v.y);
                }
            }
            if (mask & ((1u << CPQ_NUM_BANDS) - 1u))
            {
#pragma unroll
                for (int j = 0; j < kEqL / 2; ++j)
                {
                    const double2 v = reinterpret_cast<const double2*>(myStash)[j];
                    x[2 * j] = v.x + acc[2 * j];
                    x[2 * j + 1] = v.y + acc[2 * j + 1];
                }
            }
            exactMode = false;
        }
        else if (!exactMode)
        {
            for (int b = 0; b < CPQ_NUM_BANDS; ++b)
            {
                if (!((mask >> b) & 1u)) continue;   // uniform per CTA
                const double* __restrict__ bc = cst + b * kStr;
                double ic1, ic2;
                if (bandStart(std::integral_constant<int, kEqBailOut> {}, b, bc, true, ic1, ic2, bc[6] != 0.0))
                {
                    exactMode = true;   // nothing of this band has been posted: exact mode replays the bands before it and runs it serially
                    break;
                }
                linked = b;
                unsigned hiMax = max((unsigned) __double2hiint(ic1) & 0x7fffffffu, (unsigned) __double2hiint(ic2) & 0x7fffffffu);
                bool rare = (hiMax >= 0x426d1a94u) | (bc[6] != 0.0);   // |state| >= 1e12, or coefficients outside the fast path's contract
                hiMax = 0;
                const int kind = (int) bc[7];
#if CPQ_EQ_ILP2
                if (sat > 0.0)
                {
                    if (kind == 1) eq_pass2x2<true, 1>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                    else if (kind == 2) eq_pass2x2<true, 2>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                    else eq_pass2x2<true, 0>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                }
                else
                {
                    if (kind == 1) eq_pass2x2<false, 1>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                    else if (kind == 2) eq_pass2x2<false, 2>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                    else eq_pass2x2<false, 0>(x, ic1, ic2, mid1, mid2, bc, alpha, gamma, hiMax);
                }
#else
                if (sat > 0.0)
                {
                    if (kind == 1) eq_pass2<true, 1>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                    else if (kind == 2) eq_pass2<true, 2>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                    else eq_pass2<true, 0>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                }
                else
                {
                    if (kind == 1) eq_pass2<false, 1>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                    else if (kind == 2) eq_pass2<false, 2>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                    else eq_pass2<false, 0>(x, ic1, ic2, bc, alpha, gamma, hiMax);
                }
#endif
                rare |= hiMax >= thrHi;   // some |out| >= 4.5 (100 without saturation), or NaN
                if (__any_sync(0xffffffffu, rare))
                {
                    exactMode = true;
                    break;
                }
            }
        }
        // ---- exact mode: the whole segment again from the tile input with the reference's per-sample semantics; bands
        // whose link was already served are replayed (same states up to rounding), the others take part in the chain ----
        if (exactMode)
        {
            unsigned dummy = 0;
            loadBlock(dummy);
            bool firstBand = true;   // the band that filters the raw input
            for (int b = 0; b < CPQ_NUM_BANDS; ++b)
            {
                if (!((mask >> b) & 1u)) continue;
                const double* __restrict__ bc = cst + b * kStr;
                double ic1, ic2;
                const bool ranSerial = bandStart(std::integral_constant<int, kEqSerial> {}, b, bc, b > linked, ic1, ic2, (rawBig && firstBand) || bc[6] != 0.0);
                firstBand = false;
                // a reset here was not foreseen (look-back links with a start state above 1e9): never different numbers
                if (!ranSerial && eq_pass2_exact(x, ic1, ic2, bc, sat, (scalarBits >> b) & 1u)) atomicExch(a.fault, 1u);
            }
        }
        // ---- linear output stages (kept out of the band loop so that the hot loop's code is not disturbed) ----
        if (postMask)
        {
            bool gainDone = false, makeupDone = false;
            for (int b = CPQ_NUM_BANDS; b < kEqStages; ++b)
            {
                if (!((mask >> b) & 1u)) continue;
                const double* __restrict__ bc = cstPost + (b - CPQ_NUM_BANDS) * kStr;
                double ic1, ic2;
                preStage(b, gainDone, makeupDone);
                bandStart(std::integral_constant<int, kEqNoSerial> {}, b, bc, true, ic1, ic2, false);
                postPass2(bc, ic1, ic2);
            }
        }
    }
    if (STATS && a.sumsqOut) blockStats(a.sumsqOut);   // a STATS launch never runs output stages: x is the band output
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kEqL / 2; ++j) reinterpret_cast<double2*>(myStash)[j] = make_double2(x[2 * j], x[2 * j + 1]);
    __syncwarp();

    // ---- store: total gain ramp, makeup, headroom (three separate roundings, as the reference applies them), then the
    // optional output scrub + hard clamp of processOutputDouble (DSPCoreDouble.cpp:665-691, 712-737) ----
    {
        double* op = io + w0;
        const bool mulG = a.doGain && !gainInReg, mulM = a.doEpilogue && !makeupInReg, mulH = a.doEpilogue && a.applyHeadroom;
        const bool scrubOut = POST && a.doEpilogue && (a.finalClamp & 1), clampOut = POST && a.doEpilogue && (a.finalClamp & 2);
        bool overLim = false;   // SimplePeakLimiter is the identity while every |x| <= clipStart (its envelope stays exactly 1)
        const double gconst = (mulG && !a.gainTab) ? __ldg(a.gainConst + set) : 1.0;
        const double mk = a.makeup;
        constexpr double hr = 0.8912509381337456;
        auto finish = [&](double v) {
            if (mulM) v *= mk;
            if (mulH) v *= hr;
            if (scrubOut && !(fabs(v) < 1.0e300)) v = 0.0;
            if (POST) overLim |= !(fabs(v) <= 0.8413951287507587 - 0.108748 * 0.5);
            if (clampOut) v = fmin(fmax(v, -hr), hr);
            return v;
        };
        if (mulG && a.gainTab)
        {
            for (int i = lane; i < nValid; i += 32)
            {
                double v = wtile[eq_sidx(i)];
                const int64_t t = w0 + i;
                const int64_t c = cbOf(t);
                const int off = (int) (t - c * a.blockSize);
                const double2 g = __ldg(reinterpret_cast<const double2*>(a.gainTab) + (size_t) gainRow * a.nCallbacks + c);
                v *= fma((double) off, g.y, g.x);
                op[i] = finish(v);
            }
        }
        else if ((nValid & 1) == 0 && (reinterpret_cast<uintptr_t>(op) & 15) == 0)
        {
#pragma unroll
            for (int k = 0; k < kEqL / 2; ++k)
            {
                const int i = 2 * (lane + 32 * k);
                if (i < nValid)
                {
                    double2 v = *reinterpret_cast<const double2*>(wtile + eq_sidx(i));
                    if (mulG) { v.x *= gconst; v.y *= gconst; }
                    v.x = finish(v.x);
                    v.y = finish(v.y);
                    *reinterpret_cast<double2*>(op + i) = v;
                }
            }
        }
        else
        {
#pragma unroll
            for (int k = 0; k < kEqL; ++k)
            {
                const int i = lane + 32 * k;
                if (i < nValid)
                {
                    double v = wtile[eq_sidx(i)];
                    if (mulG) v *= gconst;
                    op[i] = finish(v);
                }
            }
        }
        if (POST && a.limFlag && __any_sync(0xffffffffu, overLim) && lane == 0) atomicOr(a.limFlag + seq / a.limDiv, 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// AGC (EQProcessor::processAGC + calculateAGCGain, EQProcessor.Processing.cpp:343-445): a block-rate scalar recurrence
// per stream over the callback statistics the EQ launch left behind; one thread per stream writes the gain ramp
// (start, increment) of every callback, which a second eq_kernel launch applies (applyGainRamp, :441-444).
// ---------------------------------------------------------------------------------------------
struct AgcArgs
{
    const double* sumsqIn;    // [nStreams * nch][nCallbacks]
    const double* sumsqOut;
    double* gainTab;          // [nStreams][nCallbacks][2]
    const uint8_t* agcOn;     // [nStreams]; streams without AGC get their settled total gain
    const double* gainConst;  // [nSets]
    const int* setOfSeq;      // [nStreams * nch]
    double* stateOut;         // nullable [nStreams][3] envIn, envOut, gain after the last callback
    const double* stateIn;    // nullable [nStreams][3]: the same at the start of the call (streaming continuation)
    int nStreams, nch;
    int64_t nCallbacks;
    double blockN;            // samples per callback
    double attack, release, smooth;   // 1 - exp(-n / (sr tau)) (EQProcessor.Core.cpp:776-785)
};

__global__ void agc_kernel(AgcArgs a)
{
    const int st = blockIdx.x * blockDim.x + threadIdx.x;
    if (st >= a.nStreams) return;
    double2* tab = reinterpret_cast<double2*>(a.gainTab) + (size_t) st * a.nCallbacks;
    if (!a.agcOn[st])
    {
        const double g = a.gainConst[a.setOfSeq[(size_t) st * a.nch]];
        for (int64_t c = 0; c < a.nCallbacks; ++c) tab[c] = make_double2(g, 0.0);
        return;
    }
    double envIn = 0.0, envOut = 0.0, cur = 1.0;   // rtAgc*Shadow after the reset prepareToPlay requests
    if (a.stateIn)
    {
        envIn = a.stateIn[(size_t) st * 3];
        envOut = a.stateIn[(size_t) st * 3 + 1];
        cur = a.stateIn[(size_t) st * 3 + 2];
    }
    for (int64_t c = 0; c < a.nCallbacks; ++c)
    {
        double inRms = 0.0, outRms = 0.0;
        for (int ch = 0; ch < a.nch; ++ch)
        {
            const size_t q = ((size_t) st * a.nch + ch) * a.nCallbacks + c;
            inRms = fmax(inRms, __dsqrt_rn(__ddiv_rn(a.sumsqIn[q], a.blockN)));     // NaN-free maximum like `if (rms > max)`
            outRms = fmax(outRms, __dsqrt_rn(__ddiv_rn(a.sumsqOut[q], a.blockN)));
        }
        if (!(fabs(inRms) <= 1000.0)) inRms = 1000.0;    // non-finite or > MAX_ENV_VALUE
        if (!(fabs(outRms) <= 1000.0)) outRms = 1000.0;
        const double aIn = inRms > envIn ? a.attack : a.release, aOut = outRms > envOut ? a.attack : a.release;
        envIn = __dadd_rn(__dmul_rn(envIn, 1.0 - aIn), __dmul_rn(inRms, aIn));
        envOut = __dadd_rn(__dmul_rn(envOut, 1.0 - aOut), __dmul_rn(outRms, aOut));
        if (envIn < 1.0e-20) envIn = 0.0;
        if (envOut < 1.0e-20) envOut = 0.0;
        double target = 1.0;
        if (!(envOut < 1.0e-6))
        {
            const double ratio = __ddiv_rn(envIn, envOut);
            if (!(ratio > 1.0 / 1.059 && ratio < 1.059)) target = fmin(fmax(ratio, (double) 0.06f), (double) 16.0f);
        }
        const double next = __dadd_rn(__dmul_rn(cur, 1.0 - a.smooth), __dmul_rn(target, a.smooth));
        tab[c] = make_double2(cur, __ddiv_rn(next - cur, a.blockN));
        cur = next;
    }
    if (a.stateOut)
    {
        a.stateOut[(size_t) st * 3] = envIn;
        a.stateOut[(size_t) st * 3 + 1] = envOut;
        a.stateOut[(size_t) st * 3 + 2] = cur;
    }
}

// ---------------------------------------------------------------------------------------------
// SimplePeakLimiter (audioengine/SimplePeakLimiter.h:36-86; processOutputDouble, DSPCoreDouble.cpp:700-710) between the
// scrub and the hard clamp: threshold kOutputHeadroom - 0.5 dB, knee 1 dB, immediate attack, exponential release, one
// envelope per stream for both channels.  The envelope update `want < env ? want : 1 + (env - 1) r` is not a monotone map
// of env, so it does not compose into a scan; but the stage is exactly the identity for a stream none of whose samples
// exceeds threshold - knee/2 (the envelope never leaves 1.0), which the EQ launch's store stage detects per stream.
// Flagged streams run the recurrence serially, one thread per stream, together with the hard clamp that follows it.
// ---------------------------------------------------------------------------------------------
struct LimiterArgs
{
    double* io;             // this chunk's [nStreams * nch][stride]
    int64_t stride, T;
    int nStreams, nch;
    const unsigned* flag;   // [nStreams]
    double release;         // exp(-1 / (sr * releaseSeconds))
    int clamp;              // +-kOutputHeadroom after the limiter
    double* envOut;         // nullable [nStreams]
    const double* envIn;    // nullable [nStreams]: envelope at the start of the call (streaming continuation); null = 1.0
};

__global__ void limiter_kernel(LimiterArgs a)
{
    const int st = blockIdx.x * blockDim.x + threadIdx.x;
    if (st >= a.nStreams) return;
    const double env0 = a.envIn ? a.envIn[st] : 1.0;
    if (!a.flag[st] && env0 == 1.0)   // identity: no sample can engage the limiter and the envelope rests at 1
    {
        if (a.envOut) a.envOut[st] = 1.0;
        return;
    }
    constexpr double thr = 0.8413951287507587, knee = 0.108748, clipStart = thr - knee * 0.5, hr = 0.8912509381337456;
    double* L = a.io + (size_t) st * a.nch * a.stride;
    double* R = a.nch > 1 ? L + a.stride : nullptr;
    double env = env0;
    for (int64_t i = 0; i < a.T; i += 2)   // T is even (a multiple of the block)
    {
        double2 l = *reinterpret_cast<const double2*>(L + i);
        double2 r = R ? *reinterpret_cast<const double2*>(R + i) : l;
        double lv[2] = { l.x, l.y }, rv[2] = { r.x, r.y };
#pragma unroll
        for (int k = 0; k < 2; ++k)
        {
            const double al = fabs(lv[k]), ar = R ? fabs(rv[k]) : al;
            const double peak = al < ar ? ar : al;
            const double sp = peak < 1.0e-12 ? 1.0e-12 : peak;
            double want = 1.0;
            if (sp > clipStart)
            {
                if (sp <= thr)
                {
                    const double t = __ddiv_rn(sp - clipStart, knee);
                    const double shape = __dmul_rn(__dmul_rn(t, t), __dadd_rn(3.0, -__dmul_rn(2.0, t)));
                    want = __dadd_rn(1.0, -__dmul_rn(__dadd_rn(1.0, -__ddiv_rn(thr, sp)), shape));
                }
                else want = __ddiv_rn(thr, sp);
            }
            env = want < env ? want : __dadd_rn(1.0, __dmul_rn(env - 1.0, a.release));
            lv[k] *= env;
            rv[k] *= env;
            if (a.clamp)
            {
                lv[k] = fmin(fmax(lv[k], -hr), hr);
                rv[k] = fmin(fmax(rv[k], -hr), hr);
            }
        }
        *reinterpret_cast<double2*>(L + i) = make_double2(lv[0], lv[1]);
        if (R) *reinterpret_cast<double2*>(R + i) = make_double2(rv[0], rv[1]);
    }
    if (a.envOut) a.envOut[st] = env;
}

// ---------------------------------------------------------------------------------------------
// Float host buffers (the application's float path converts on entry and exit: convertFloatToDoubleHighQuality's cast,
// InputBitDepthTransform.h:102-121, and static_cast<float> of the clamped result, AudioEngine.Processing.DSPCoreIO.cpp:524-537):
// the wire format is FP32, the arithmetic stays FP64.
// ---------------------------------------------------------------------------------------------
template <bool TO_DOUBLE>
__global__ void convert_kernel(float* __restrict__ f, double* __restrict__ d, int64_t dStride, int64_t T)
{
    // pairs: float rows have the pitch T rounded up to even, so an odd T carries one pad sample through like the double rows do
    float2* f2 = reinterpret_cast<float2*>(f + (size_t) blockIdx.y * ((T + 1) & ~(int64_t) 1));
    double2* d2 = reinterpret_cast<double2*>(d + (size_t) blockIdx.y * dStride);
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < (T + 1) / 2; i += (int64_t) gridDim.x * blockDim.x)
    {
        if (TO_DOUBLE)
        {
            const float2 v = f2[i];
            d2[i] = make_double2((double) v.x, (double) v.y);
        }
        else
        {
            const double2 v = d2[i];
            f2[i] = make_float2(__double2float_rn(v.x), __double2float_rn(v.y));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The engine's input stage (convo::input_transform::applyHighQuality64BitTransform, InputBitDepthTransform.h:86-100, from
// DSPCore::processInput): optional gain, NaN or |v| < 1e-20 -> 0, clamp to [-1, 1] (+-Inf survives the scrub of the
// reference's four-wide body and clamps to +-1; only its scalar remainder, n % 4 samples per callback, zeroes it).
// ---------------------------------------------------------------------------------------------
struct InputArgs
{
    double* io;
    int64_t stride, T;
    double gain;
    int applyGain;
    int block, vecEnd;   // samples per callback and block / 4 * 4: offsets >= vecEnd take the scalar remainder's rule
};

__global__ void input_kernel(InputArgs a)
{
    double* io = a.io + (size_t) blockIdx.y * a.stride;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < a.T; i += (int64_t) gridDim.x * blockDim.x)
    {
        double v = io[i];
        if (a.applyGain) v *= a.gain;
        const bool tailSample = (int) (i % a.block) >= a.vecEnd;
        if (v != v || fabs(v) < 1.0e-20 || (tailSample && isinf(v))) v = 0.0;
        io[i] = fmin(fmax(v, -1.0), 1.0);
    }
}

// ---------------------------------------------------------------------------------------------
// Dry path of ConvolverProcessor::process in its settled state (ConvolverProcessor.Runtime.cpp:551-568, 573-584, 675-677,
// 748): io already holds scrub(wet) * wetGain (eq_kernel's load stage); add the input delayed by the latency-compensation
// delay times equalPowerSin(1 - mix), or -- dry-only fast path, mix <= 0.001 -- replace io by the delayed input.
// ---------------------------------------------------------------------------------------------
struct MixArgs
{
    double* io;          // [nSeq][stride]
    const double* dry;   // [nSeq][stride] copy of the convolver input
    int64_t stride, T;
    int delay;           // samples; the delay ring starts zeroed
    double dryGain;
    int dryOnly;
    // streaming continuation: the last `delay` samples of the convolver input before this call (the delay ring's content),
    // [nSeq][delay], and where the ring's content after this call goes (another buffer); both nullable
    const double* hist;
    double* histOut;
};

__global__ void mix_kernel(MixArgs a)
{
    double* io = a.io + (size_t) blockIdx.y * a.stride;
    const double* dry = a.dry + (size_t) blockIdx.y * a.stride;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < a.T; i += (int64_t) gridDim.x * blockDim.x)
    {
        const double d = i >= a.delay ? dry[i - a.delay] : (a.hist ? a.hist[(size_t) blockIdx.y * a.delay + i] : 0.0);
        io[i] = a.dryOnly ? d : __dadd_rn(io[i], __dmul_rn(d, a.dryGain));
    }
    if (a.histOut)   // the last `delay` samples of [ring | this call's input]
        for (int64_t j = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; j < a.delay; j += (int64_t) gridDim.x * blockDim.x)
        {
            const int64_t p = a.T - a.delay + j;
            a.histOut[(size_t) blockIdx.y * a.delay + j] = p >= 0 ? dry[p] : (a.hist ? a.hist[(size_t) blockIdx.y * a.delay + a.delay + p] : 0.0);
        }
}

// ---------------------------------------------------------------------------------------------
// Experimental direct-form head (MKLNonUniformConvolver::processDirectBlock, MKLNonUniformConvolver.cpp:1169-1232; added
// to the L0 output in Get, :1604-1616): io += FIR of the first <= 32 taps over the convolver input, accumulated like the
// reference (two 4-lane FMA accumulators over reversed taps, lanes summed (v0 + v2) + (v1 + v3)); |y| < 1e-20 or
// non-finite -> 0.  Those taps are zeroed in the impulse the partitions were built from (:730-731) and bypass the
// spectrum filter, so with a FilterSpec the head changes the result beyond rounding.
// ---------------------------------------------------------------------------------------------
struct DirectArgs
{
    double* io;            // [nSeq][stride]: L0 output so far
    const double* x;       // [nSeq][stride]: the convolver input (copy)
    const double* taps;    // [nH][32]: hrev[k] = h[31 - k] * scale, zero padded at the front for shorter heads
    int64_t stride, T;
    const double* histEnd; // nullable (Reset state before the call): histEnd[seq * histStride + g] = input sample g < 0 (streaming continuation)
    int64_t histStride;
    int hSeqMod, seqBase;  // row = (seqBase + seq) % hSeqMod when the IR pair is shared; 0 = seqBase + seq
};

__global__ void direct_head_kernel(DirectArgs a)
{
    __shared__ double h[32];
    const int seq = blockIdx.y;
    const int row = a.hSeqMod > 0 ? (a.seqBase + seq) % a.hSeqMod : a.seqBase + seq;
    if (threadIdx.x < 32) h[threadIdx.x] = a.taps[(size_t) row * 32 + threadIdx.x];
    __syncthreads();
    double* io = a.io + (size_t) seq * a.stride;
    const double* x = a.x + (size_t) seq * a.stride;
    const double* hist = a.histEnd ? a.histEnd + (size_t) seq * a.histStride : nullptr;
    for (int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < a.T; t += (int64_t) gridDim.x * blockDim.x)
    {
        double s[8] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
        for (int k = 0; k < 32; ++k)
        {
            const int64_t src = t - 31 + k;
            const double v = src >= 0 ? x[src] : (hist ? hist[src] : 0.0);
            s[k & 7] = fma(h[k], v, s[k & 7]);
        }
        double y = __dadd_rn(__dadd_rn(__dadd_rn(s[0], s[4]), __dadd_rn(s[2], s[6])), __dadd_rn(__dadd_rn(s[1], s[5]), __dadd_rn(s[3], s[7])));
        if (!(fabs(y) >= 1.0e-20 && fabs(y) <= 1.79769313486231570815e308)) y = 0.0;
        io[t] += y;
    }
}

// ---------------------------------------------------------------------------------------------
// Mid/Side bands (node path, EQProcessor.Processing.cpp:690-740): encode (L+R)/2, (L-R)/2 of the listed streams into a
// scratch pair of rows, run the band on one of them with eq_kernel, decode L = M+S, R = M-S.
// ---------------------------------------------------------------------------------------------
struct MsArgs
{
    double* io;          // this chunk's [nSeq][ioStride]
    int64_t ioStride;
    double* ms;          // [2 * nStreams][ioStride]: Mid row, Side row per listed stream
    const int* streams;  // [nStreams] absolute stream indices
    int streamBase;      // first stream of the chunk
    int64_t T;
};

// MODE 0 encode (M, S) <- (L, R);  1 decode (L, R) <- (M, S);  and for Mid/Side bands inside the Parallel structure, where the
// rows carry the bands' summed differences:  2 rows -= encode(L, R);  3 (L, R) += (M + S, M - S).
template <int MODE>
__global__ void ms_kernel(MsArgs a)
{
    const int st = a.streams[blockIdx.y] - a.streamBase;
    double2* L = reinterpret_cast<double2*>(a.io + (size_t) (2 * st) * a.ioStride);
    double2* R = reinterpret_cast<double2*>(a.io + (size_t) (2 * st + 1) * a.ioStride);
    double2* M = reinterpret_cast<double2*>(a.ms + (size_t) (2 * blockIdx.y) * a.ioStride);
    double2* S = reinterpret_cast<double2*>(a.ms + (size_t) (2 * blockIdx.y + 1) * a.ioStride);
    const int64_t n2 = (a.T + 1) / 2;   // an odd T takes the row's pad sample along
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t) gridDim.x * blockDim.x)
    {
        if (MODE == 0 || MODE == 2)
        {
            const double2 l = L[i], r = R[i];
            const double2 m = make_double2((l.x + r.x) * 0.5, (l.y + r.y) * 0.5);
            const double2 sd = make_double2((l.x - r.x) * 0.5, (l.y - r.y) * 0.5);
            if (MODE == 0)
            {
                M[i] = m;
                S[i] = sd;
            }
            else
            {
                const double2 pm = M[i], ps = S[i];
                M[i] = make_double2(pm.x - m.x, pm.y - m.y);
                S[i] = make_double2(ps.x - sd.x, ps.y - sd.y);
            }
        }
        else
        {
            const double2 m = M[i], sd = S[i];
            if (MODE == 1)
            {
                L[i] = make_double2(m.x + sd.x, m.y + sd.y);
                R[i] = make_double2(m.x - sd.x, m.y - sd.y);
            }
            else
            {
                const double2 l = L[i], r = R[i];
                L[i] = make_double2(l.x + (m.x + sd.x), l.y + (m.y + sd.y));
                R[i] = make_double2(r.x + (m.x - sd.x), r.y + (m.y - sd.y));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Dither + 12-tap error-feedback noise shaper (PsychoacousticDither::processStereoBlock, PsychoacousticDither.h:293-405).
//
// Serial in time by nature: e[n] feeds tmp[n+1] through the quantiser, and the recurrence is chaotic (the feedback gains reach
// 11.7, so a one-ulp difference flips the quantiser within a few hundred samples).  Parity is therefore bit-for-bit or nothing,
// and the arithmetic below is the association of the reference as compiled by oracle/Makefile (g++ -O2 -mfma contracts
// a*b + c): shaped = fma(c11,z11, ... fma(c2,z2, fma(c0,z0, c1*z1))); left channel of a stereo stream
// tmp = fma(x, headroom, tpdf*scale) + shaped, right channel / mono tmp = fma(tpdf, scale, x*headroom) + shaped.
// The dependent chain from e[n-1] to e[n] is 16 FP64 operations (11 FMA, add, mul, round, mul, sub) = about 130 cycles per
// sample whatever the batch size -- the floor of this stage is T x 130 cycles (33 ms for 480 256 samples at 1.9 GHz), so the
// engine runs it on a side stream where it overlaps the next chunk's transforms.
//
// One lane per sequence, a warp = 32 sequences.  Memory goes through shared memory in [32 sequences][12 samples] tiles (the
// one-thread-per-sequence form touched 32 sectors per warp load): cp.async brings tile k+1 (signal + 2 uniforms per sample)
// while tile k is computed; results are written back from shared memory as 16-byte stores.  A tile is twelve samples because
// that is the order of the shaper: the loop over a tile is fully unrolled and the error history rotates through twelve named
// registers, one turn per tile, instead of being shifted by eleven moves per sample.
// ---------------------------------------------------------------------------------------------
struct DitherArgs
{
    double* io;
    int64_t ioStride;
    int64_t T;
    int nSeq;
    int nch;                  // channels per stream: lane role = left channel of a stereo stream or not
    int seqBase;              // absolute index of this launch's first sequence (a chunk may start on a right channel)
    const double* uniforms;   // [nSeq][2*T]; nullptr = generate them: the reference's own fallback generator (rng below)
    unsigned long long* rng;  // [nSeq] xorshift64* state per stream-channel (PsychoacousticDither::fallbackState, :485-497), carried
    double coeff[12];
    double scale, invScale;
    int64_t uniRow;           // samples per channel in a row of `uniforms` (= T unless the call is one time segment of a longer buffer)
    double* z;                // [nSeq][12] error history (carried)
    int finalClamp;           // after the quantiser: bit 0 scrub, bit 1 clamp to +-kOutputHeadroom (DSPCoreDouble.cpp:665-691, 712-737)
};

constexpr int kDthTile = 24;                 // samples per tile and sequence: a multiple of the order of the shaper, so the 12-deep error
                                             // history rotates through registers by name, two full turns per tile (no moves)
constexpr int kDthStages = 4;                // tile k is computed while k + 1 .. k + 3 are in flight / being written back
constexpr int kDthRow = kDthTile + 2;        // signal row pitch in doubles (16-byte aligned rows)
constexpr int kDthURow = 2 * kDthTile + 2;   // uniform row pitch
constexpr int kDthBufDoubles = 32 * kDthRow + 32 * kDthURow;
constexpr size_t kDitherSmemBytes = (size_t) kDthStages * kDthBufDoubles * sizeof(double) + 2 * kDthStages * sizeof(uint64_t);
constexpr int kDitherThreads = 64;           // warp 0 runs the shaper, warp 1 moves the tiles

__device__ __forceinline__ void dthCpAsync16(void* smem, const void* gmem)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
// the mbarrier receives one arrival from this thread once all its earlier cp.async copies have landed
__device__ __forceinline__ void dthCpAsyncArrive(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// one sample: e[] is the error history, logical z[t] = e[(t + ROT) % 12]; the new error replaces the oldest entry
template <int ROT>
__device__ __forceinline__ double dither_sample(double x, double2 uu, double (&e)[12], const double (&c)[12], bool roleLeft, double scale,
                                                double invScale, int finalClamp)
{
    constexpr double kHeadroom = 0.8912509381337456;
    double shaped = __dmul_rn(c[1], e[(1 + ROT) % 12]);
    shaped = __fma_rn(c[0], e[ROT % 12], shaped);
#pragma unroll
    for (int t = 2; t < 12; ++t) shaped = __fma_rn(c[t], e[(t + ROT) % 12], shaped);
    const double tpdf = __dadd_rn(__dadd_rn(uu.x, -0.5), __dadd_rn(uu.y, -0.5));
    const double pre = roleLeft ? __fma_rn(x, kHeadroom, __dmul_rn(tpdf, scale)) : __fma_rn(tpdf, scale, __dmul_rn(x, kHeadroom));
    const double tmp = __dadd_rn(pre, shaped);
    const double q = __dmul_rn(rint(__dmul_rn(tmp, invScale)), scale);
    double err = __dadd_rn(tmp, -q);
    if (fabs(err) < 1.0e-20) err = 0.0;
    e[(11 + ROT) % 12] = err;   // the oldest entry (logical z[11]) becomes the newest of the next sample (ROT - 1)
    double o = q;
    if ((finalClamp & 1) && !(fabs(o) < 1.0e300)) o = 0.0;
    if (finalClamp & 2) o = fmin(fmax(o, -kHeadroom), kHeadroom);
    return o;
}

// PsychoacousticDither::fallbackUniform (:485-497): xorshift64* step, top 53 bits -> [0, 1)
__device__ __forceinline__ double dither_fallback_uniform(unsigned long long& x)
{
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    const unsigned long long z = x * 2685821657736338717ull;
    return (double) (z >> 11) * (1.0 / 9007199254740992.0);
}

template <int J>
__device__ __forceinline__ void dither_unrolled(double* mine, const double* myU, double (&e)[12], const double (&c)[12], bool roleLeft,
                                                double scale, double invScale, int finalClamp)
{
    if constexpr (J < kDthTile)
    {
        const double2 uu = *reinterpret_cast<const double2*>(myU + 2 * J);
        mine[J] = dither_sample<(24 - J) % 12>(mine[J], uu, e, c, roleLeft, scale, invScale, finalClamp);
        dither_unrolled<J + 1>(mine, myU, e, c, roleLeft, scale, invScale, finalClamp);
    }
}

__global__ void __launch_bounds__(kDitherThreads) dither_kernel(DitherArgs a)
{
    extern __shared__ __align__(16) double dthSmem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(dthSmem + (size_t) kDthStages * kDthBufDoubles);   // "tile landed" (32 copy lanes)
    uint64_t* done = full + kDthStages;                                                               // "tile shaped" (1 arrival)
    const int lane = threadIdx.x & 31;
    const bool mover = threadIdx.x >= 32;
    const int seq0 = blockIdx.x * 32;
    const int nLocal = min(32, a.nSeq - seq0);
    const int64_t nTiles = (a.T + kDthTile - 1) / kDthTile;
    const bool useRng = a.uniforms == nullptr;
    if (threadIdx.x == 0)
        for (int i = 0; i < kDthStages; ++i)
        {
            mbar_init(full + i, useRng ? 64 : 32);   // the copy lanes' asynchronous arrivals (+ their own once the drawn uniforms are stored)
            mbar_init(done + i, 1);
        }
    __syncthreads();
    auto tileSamples = [&](int64_t tile) { const int64_t r = a.T - tile * kDthTile; return (int) (r < (int64_t) kDthTile ? r : (int64_t) kDthTile); };

    if (mover)
    {
        // ---- warp 1: tiles in (cp.async, 16-byte pieces, completion counted on the stage's mbarrier) and shaped tiles out.  With one
        // warp doing both jobs the staging loops were more instructions than the shaper and ran in series with its dependent
        // chain (304 cycles per sample against the chain's 190, profiles/r02d_stalls_dither_kernel.txt). ----
        constexpr int kSigPieces = kDthTile / 2;               // 16-byte pieces per signal row
        constexpr int rowsPerStep = 32 / kSigPieces;           // signal rows covered by one warp-wide step
        const int rr = lane / kSigPieces, pc = lane % kSigPieces;
        const size_t sigRowStep = (size_t) rowsPerStep * a.ioStride;
        double* const sigBase = a.io + (size_t) (seq0 + rr) * a.ioStride + 2 * pc;                 // + tile * kDthTile, + k * sigRowStep
        const double* const uniBase = useRng ? nullptr : a.uniforms + ((size_t) seq0 * a.uniRow + lane) * 2;   // + tile * 2 kDthTile, + r * 2 T
        const size_t uniRowStep = (size_t) a.uniRow * 2;
        // cpq_set_dither_seed: the uniforms come from the reference's fallback generator, one stream per sequence -- drawn here,
        // a tile ahead, so that the integer work stays out of the shaper's instruction stream (277 -> 207 cycles per sample)
        unsigned long long rs = (useRng && lane < nLocal) ? a.rng[seq0 + lane] : 0ull;
        auto issue = [&](int64_t tile) {
            const int buf = (int) (tile % kDthStages);
            double* sig = dthSmem + (size_t) buf * kDthBufDoubles;
            double* uni = sig + 32 * kDthRow;
            const int n = tileSamples(tile);
            if (rr < rowsPerStep && 2 * pc < n)   // an odd n (odd T) takes the row's pad sample along
            {
                const double* g = sigBase + tile * kDthTile;
                double* d = sig + rr * kDthRow + 2 * pc;
                for (int r = rr; r < nLocal; r += rowsPerStep, g += sigRowStep, d += rowsPerStep * kDthRow) dthCpAsync16(d, g);
            }
            if (!useRng && lane < n)
            {
                const double* g = uniBase + tile * (2 * kDthTile);
                double* d = uni + 2 * lane;
                for (int r = 0; r < nLocal; ++r, g += uniRowStep, d += kDthURow) dthCpAsync16(d, g);
            }
            dthCpAsyncArrive(full + buf);
            if (useRng)
            {
                if (lane < nLocal)
                {
                    double* d = uni + lane * kDthURow;
                    for (int i = 0; i < 2 * n; ++i) d[i] = dither_fallback_uniform(rs);   // u1 then u2, as nextTPDF_MKL draws them (:560-572)
                }
                mbar_arrive(full + buf);   // release: the shaper's wait sees the stores
            }
        };
        for (int64_t t = 0; t < kDthStages && t < nTiles; ++t) issue(t);
        for (int64_t tile = 0; tile < nTiles; ++tile)
        {
            const int buf = (int) (tile % kDthStages);
            mbar_wait(done + buf, (unsigned) ((tile / kDthStages) & 1));
            const int n = tileSamples(tile);
            if (rr < rowsPerStep && 2 * pc < n)
            {
                const double* sSrc = dthSmem + (size_t) buf * kDthBufDoubles + rr * kDthRow + 2 * pc;
                double* g = sigBase + tile * kDthTile;
                for (int r = rr; r < nLocal; r += rowsPerStep, g += sigRowStep, sSrc += rowsPerStep * kDthRow)
                    *reinterpret_cast<double2*>(g) = *reinterpret_cast<const double2*>(sSrc);
            }
            __syncwarp();   // every lane has read its pieces before the stage is refilled
            if (tile + kDthStages < nTiles) issue(tile + kDthStages);
        }
        if (useRng && lane < nLocal) a.rng[seq0 + lane] = rs;
        return;
    }

    // ---- warp 0: the shaper, one lane per sequence ----
    const int seq = seq0 + min(lane, nLocal - 1);          // idle lanes shadow the last sequence, they never write
    const bool live = lane < nLocal;
    const bool roleLeft = a.nch == 2 && ((a.seqBase + seq) % 2) == 0;
    double e[12];   // e[t] = logical z[t] between tiles
#pragma unroll
    for (int i = 0; i < 12; ++i) e[i] = a.z[(size_t) seq * 12 + i];
    double c[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) c[i] = a.coeff[i];
    for (int64_t tile = 0; tile < nTiles; ++tile)
    {
        const int buf = (int) (tile % kDthStages);
        mbar_wait(full + buf, (unsigned) ((tile / kDthStages) & 1));
        double* sig = dthSmem + (size_t) buf * kDthBufDoubles;
        const double* uni = sig + 32 * kDthRow;
        const int n = tileSamples(tile);
        double* mine = sig + lane * kDthRow;
        const double* myU = uni + lane * kDthURow;
        if (live)
        {
            if (n == kDthTile) dither_unrolled<0>(mine, myU, e, c, roleLeft, a.scale, a.invScale, a.finalClamp);
            else
                for (int i = 0; i < n; ++i)   // last, partial tile: shift the history like the reference does
                {
                    const double2 uu = *reinterpret_cast<const double2*>(myU + 2 * i);
                    mine[i] = dither_sample<0>(mine[i], uu, e, c, roleLeft, a.scale, a.invScale, a.finalClamp);
                    const double newest = e[11];
#pragma unroll
                    for (int t = 11; t > 0; --t) e[t] = e[t - 1];
                    e[0] = newest;
                }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(done + buf);   // release: the mover's wait sees every lane's samples
    }
    if (live)
    {
#pragma unroll
        for (int i = 0; i < 12; ++i) a.z[(size_t) seq * 12 + i] = e[i];
    }
}

// DFMA throughput probe (roofline denominator for the FP64 pipe); 8 independent chains per thread.
__global__ void dfma_probe_kernel(double* out, int iters)
{
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double m = 1.0000001, c = 1e-12;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// Dependent-issue latency of DFMA: one warp, one chain.
__global__ void dfma_latency_kernel(double* out, long long* cycles, int iters)
{
    double a = 1.0 + 1e-9 * threadIdx.x;
    const double m = 1.0000001, c = 1e-12;
    const long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < iters; ++it) a = fma(a, m, c);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    if (a == 123.456) out[0] = a;
}

} // namespace cpq
