// cpq_eq.cuh -- layer assembly + 20-band TPT-SVF EQ + gain/headroom epilogue, and the dither kernel (sm_100a).
//
//   eq_kernel      Get()'s layer sum (MKLNonUniformConvolver.cpp:1553-1634) -> optional ConvolverProcessor wet
//                  scrub/gain (ConvolverProcessor.Runtime.cpp:50-60,675,748) -> 20 x processBandStereo
//                  (EQProcessor.Processing.cpp:191-276) -> total-gain ramp (:1262-1274) -> makeup gain +
//                  kOutputHeadroom (AudioEngine.Processing.DSPCoreDouble.cpp:465-469,:655-663)
//   dither_kernel  PsychoacousticDither::processStereoBlock (PsychoacousticDither.h:293-355)
//
// The band recurrence is linear in its 2-element state (the tanh saturation only touches the band *output*,
// Processing.cpp:148-168), so each band is a blocked scan: every thread owns 16 consecutive samples in
// registers, computes the zero-state response of its block with 16 precomputed weight vectors (pass 1), a
// warp-shuffle + cross-warp prefix composes the per-block affine maps (all blocks share A^16), and pass 2
// re-runs the reference recurrence from the exact start state.  Bands are processed in order on the same
// registers, so a 4096-sample tile crosses HBM once for all 20 bands.  Across tiles the state is carried
// either inside the CTA (one CTA per sequence when the batch fills the GPU) or through per-(sequence, tile,
// band) records with release/acquire flags, tiles taking their index from an atomic ticket so that a
// predecessor is always already running.
//
// Fast path / exact path: with |out| < 4.5 before saturation the reference's clamps and scrubs are
// identities, so pass 2 runs without them and ORs a per-thread flag; a flagged thread replays its block
// from the stashed inputs with the reference's full per-sample semantics.  If a *state* leaves the
// reference's valid range (|ic| >= 1e15 or non-finite, where the reference zeroes it, :174-175/:257-258) the
// scan's linearity assumption is void and the kernel raises `fault`; the host reports CPQ_ERR_UNSUPPORTED.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/cpq.h"

namespace cpq
{

constexpr int kEqThreads = 256;
constexpr int kEqL = 16;                       // samples per thread
constexpr int kEqTile = kEqThreads * kEqL;     // 4096
constexpr int kEqWarps = kEqThreads / 32;

// per (parameter set, band) constants, all double; built by buildBandConstants() in cpq_engine.cu
constexpr int kEqcCoef = 0;      // a1,a2,a3,m0,m1,m2, forceExact(!=0), peakingPattern(!=0)
constexpr int kEqcW = 8;         // w[16][2]   zero-state weights, c = sum_j w[j] * v0[j]
constexpr int kEqcMs = 40;       // Ms[5][4]   A^(16*2^d), row-major 2x2
constexpr int kEqcMw = 60;       // A^512
constexpr int kEqcMt = 64;       // A^4096
constexpr int kEqcStride = 68;   // doubles per band (all 20 bands of a set = 10.9 KB, staged in shared memory)

struct EqChain
{
    double* rec;          // [(seq*nRuns + run)*20 + band] x {s1, s2, epoch-as-u64, pad}
    unsigned* ticket;     // CTA ticket counter
    unsigned long long epoch;
};

struct EqArgs
{
    double* io;             // [nSeq][ioStride] in/out (in place); holds y0 (or the raw input when !assemble)
    int64_t ioStride;
    int64_t T;              // samples per sequence
    int nSeq;
    int nTiles;             // ceil(T / 4096)
    int tilesPerRun;        // 1 (chained) or nTiles (one CTA per sequence)
    int nRuns;
    // assembly
    int assemble;           // add tails / apply the outer boundary
    int nTail;              // number of tail layers (0..2)
    const double* tail[2];  // [nSeq][tailStride[l]] layer output streams
    int64_t tailStride[2];
    const int64_t* tailSrc[2];     // [nCallbacks] stream position or -1
    const int32_t* blockMap[2];    // nullable: stream block -> frame
    int tailPartLog2[2];
    double tailGain[2];
    int blockLog2;          // log2(block size)
    int outer;              // CPQ_CONV_OUTER: scrub + wet gain
    double wetGain;
    // EQ
    int doEq;
    const double* eqc;      // [nSets][20][kEqcStride]
    const unsigned* bandMask;  // [nSeq] bit b = band b processed for this sequence
    const int* setOfSeq;    // [nSeq]
    const double* sat;      // [nSets]
    double* stateOut;       // [nSeq][20][2] final states
    const double* gainTab;  // nullable [nSets][nCallbacks][2] (start, inc)
    const double* gainConst;// [nSets] settled total gain (used when gainTab == nullptr)
    int64_t nCallbacks;
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

// num/den to < 1 ulp for den in [27, 210]: hardware reciprocal seed (~2^-20), one Newton step (~2^-40),
// then a residual correction of the quotient (~2^-80 before rounding).  5 FP64-pipe instructions.
__device__ __forceinline__ double div_nr(double num, double den)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    const double e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    const double q = num * r;
    const double rem = fma(-den, q, num);
    return fma(rem, r, q);
}

__device__ __forceinline__ bool eq_valid(double v) { return fabs(v) < 1.0e15; }   // false for NaN / Inf too

// padded shared index: 17-double stride per 16 samples keeps both the coalesced pass (consecutive t) and
// the per-thread pass (16 consecutive samples per lane) free of bank conflicts
__device__ __forceinline__ int eq_sidx(int t) { return t + (t >> 4); }

// The reference's per-sample semantics (processBandStereo) for one thread's block, run from shared memory.
// Used only by threads whose fast pass saw |out| >= 4.5 / suspicious state.  Returns true if a state had to be
// reset (which the scan cannot represent).
__device__ __noinline__ bool eq_exact_block(double* s /* smem, stride-1 over 16 padded slots */, double& ic1, double& ic2,
                                            double a1, double a2, double a3, double m0, double m1, double m2, double sat)
{
    bool reset = false;
    const double oneMinusSat = 1.0 - sat;
    for (int j = 0; j < kEqL; ++j)
    {
        const double v0 = s[j];
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
            const double th = (xc * (27.0 + x2)) / fma(9.0, x2, 27.0);
            out = out * oneMinusSat + th * sat;
        }
        if (!eq_valid(out)) out = 0.0;
        if (!eq_valid(ic1)) { ic1 = 0.0; reset = true; }
        if (!eq_valid(ic2)) { ic2 = 0.0; reset = true; }
        out = (out > -100.0) ? out : -100.0;
        out = (out < 100.0) ? out : 100.0;
        s[j] = out;
    }
    return reset;
}

__device__ __forceinline__ void matvec2(const double* __restrict__ m, double& p1, double& p2, double add1, double add2)
{
    const double n1 = fma(m[0], p1, fma(m[1], p2, add1));
    const double n2 = fma(m[2], p1, fma(m[3], p2, add2));
    p1 = n1;
    p2 = n2;
}

// Pass 2 (fast path) over one thread's block.  SAT: fused saturation
//   out*(1-s) + tanh27/9(out)*s  ==  out * (alpha + gamma / (out^2 + 3)),  alpha = (9-8s)/9, gamma = 8s/3
// with 1/(out^2+3) from the hardware seed refined to ~2^-60 (r0 (1 + e + e^2)); the correction term is <= 18 % of the
// result, so even the seed's 2^-20 would leave < 1e-12.  PEAK: bands with m0 == 1, m2 == 0 (every Peaking band) need
// only out = v0 + m1 v1.
template <bool SAT, bool PEAK>
__device__ __forceinline__ void eq_pass2(double (&x)[kEqL], double& ic1, double& ic2, double a1, double a2, double a3, double m0,
                                         double m1, double m2, double alpha, double gamma, unsigned& hiMax)
{
#pragma unroll
    for (int j = 0; j < kEqL; ++j)
    {
        const double v0 = x[j];
        const double v3 = v0 - ic2;
        const double v1 = fma(a1, ic1, a2 * v3);
        const double v2 = fma(a2, ic1, fma(a3, v3, ic2));
        ic1 = fma(2.0, v1, -ic1);
        ic2 = fma(2.0, v2, -ic2);
        const double out = PEAK ? fma(m1, v1, v0) : fma(m0, v0, fma(m1, v1, m2 * v2));
        hiMax = max(hiMax, (unsigned) __double2hiint(out) & 0x7fffffffu);
        if (SAT)
        {
            const double d = fma(out, out, 3.0);
            double r0;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
            const double e = fma(-d, r0, 1.0);
            const double r = fma(r0, fma(e, e, e), r0);
            x[j] = out * fma(gamma, r, alpha);
        }
        else x[j] = out;
    }
}

#ifndef CPQ_EQ_MINBLOCKS
#define CPQ_EQ_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(kEqThreads, CPQ_EQ_MINBLOCKS) eq_kernel(EqArgs a)
{
    __shared__ __align__(16) double tile[kEqTile + kEqTile / 16];
    __shared__ __align__(16) double cst[CPQ_NUM_BANDS * kEqcStride];   // this sequence's band constants
    __shared__ double warpAggBuf[2][kEqWarps][2];   // double-buffered by band parity
    __shared__ double sIn[2];
    __shared__ double carry[2][CPQ_NUM_BANDS][2];   // double-buffered by tile parity (one-CTA-per-sequence mode)
    __shared__ unsigned sTicket;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sTicket = atomicAdd(a.chain.ticket, 1u);
    __syncthreads();
    const unsigned ticket = sTicket;
    // run-major ticket order: every predecessor (same sequence, previous run) holds a smaller ticket
    const int run = (int) (ticket / (unsigned) a.nSeq);
    const int seq = (int) (ticket % (unsigned) a.nSeq);
    if (run >= a.nRuns) return;

    double* io = a.io + (size_t) seq * a.ioStride;
    const int set = a.doEq ? a.setOfSeq[seq] : 0;
    const unsigned mask = a.doEq ? a.bandMask[seq] : 0u;
    const double sat = a.doEq ? a.sat[set] : 0.0;
    const double alpha = fma(-8.0, sat, 9.0) / 9.0, gamma = 8.0 * sat / 3.0;
    const unsigned thrHi = sat > 0.0 ? 0x40120000u : 0x40590000u;   // high words of 4.5 / 100.0
    const bool chained = a.tilesPerRun == 1 && a.nRuns > 1;
    const int bmask = (1 << a.blockLog2) - 1;
    double* myStash = tile + eq_sidx(tid * kEqL);   // 16 consecutive slots (no pad boundary inside a block)

    if (a.doEq)
    {
        const double* __restrict__ src = a.eqc + (size_t) set * CPQ_NUM_BANDS * kEqcStride;
        for (int i = tid; i < CPQ_NUM_BANDS * kEqcStride / 2; i += kEqThreads)
            reinterpret_cast<double2*>(cst)[i] = __ldg(reinterpret_cast<const double2*>(src) + i);
    }
    if (tid < 2 * CPQ_NUM_BANDS * 2) (&carry[0][0][0])[tid] = 0.0;
    __syncthreads();

    for (int tl = 0; tl < a.tilesPerRun; ++tl)
    {
        const int tileIdx = run * a.tilesPerRun + tl;
        if (tileIdx >= a.nTiles) break;
        const int tpar = tl & 1;
        const int64_t t0 = (int64_t) tileIdx * kEqTile;
        const int nValid = (int) min((int64_t) kEqTile, a.T - t0);

        // ---- coalesced load + layer assembly (Get) ----
#pragma unroll 4
        for (int i = tid; i < kEqTile; i += kEqThreads)
        {
            double v = 0.0;
            if (i < nValid)
            {
                const int64_t t = t0 + i;
                v = io[t];
                if (a.assemble)
                {
                    const int64_t c = t >> a.blockLog2;
                    const int off = (int) t & bmask;
                    for (int l = 0; l < a.nTail; ++l)
                    {
                        const int64_t s = __ldg(a.tailSrc[l] + c);
                        if (s >= 0)
                        {
                            int64_t pos = s + off;
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
            }
            tile[eq_sidx(i)] = v;
        }
        __syncthreads();

        double x[kEqL];
        unsigned hiMax = 0;   // running max of |x|'s high word: raw input large enough that a state could reach 1e15?
#pragma unroll
        for (int j = 0; j < kEqL; ++j)
        {
            x[j] = myStash[j];
            hiMax = max(hiMax, (unsigned) __double2hiint(x[j]) & 0x7fffffffu);
        }
        bool suspicious = hiMax >= 0x41cdcd65u;   // |x| >= 1e9 (or NaN/Inf)

        if (a.doEq)
        {
            int parity = 0;
            for (int b = 0; b < CPQ_NUM_BANDS; ++b)
            {
                if (!((mask >> b) & 1u)) continue;   // uniform per CTA
                double (*warpAgg)[2] = warpAggBuf[parity];
                parity ^= 1;
                const double* __restrict__ bc = cst + b * kEqcStride;
                // ---- pass 1: zero-state response of this thread's 16 samples; stash the inputs ----
                double c1 = 0.0, c2 = 0.0;
#pragma unroll
                for (int j = 0; j < kEqL; ++j)
                {
                    const double2 w = reinterpret_cast<const double2*>(bc + kEqcW)[j];
                    c1 = fma(w.x, x[j], c1);
                    c2 = fma(w.y, x[j], c2);
                    myStash[j] = x[j];
                }
                // ---- warp inclusive scan of s -> A^16 s + c ----
#pragma unroll
                for (int d = 0; d < 5; ++d)
                {
                    const double p1 = __shfl_up_sync(0xffffffffu, c1, 1 << d);
                    const double p2 = __shfl_up_sync(0xffffffffu, c2, 1 << d);
                    if (lane >= (1 << d))
                    {
                        const double* m = bc + kEqcMs + 4 * d;
                        c1 = fma(m[0], p1, fma(m[1], p2, c1));
                        c2 = fma(m[2], p1, fma(m[3], p2, c2));
                    }
                }
                if (lane == 31) { warpAgg[warp][0] = c1; warpAgg[warp][1] = c2; }
                // exclusive value (state contribution before this thread, relative to the warp start)
                double e1 = __shfl_up_sync(0xffffffffu, c1, 1);
                double e2 = __shfl_up_sync(0xffffffffu, c2, 1);
                if (lane == 0) { e1 = 0.0; e2 = 0.0; }
                __syncthreads();

                double p1, p2;   // state at the start of the tile
                if (chained)
                {
                    if (tid == 0)
                    {
                        // tile aggregate with zero carry-in, then the carry-in from the previous tile's CTA
                        double g1 = 0.0, g2 = 0.0;
                        for (int w = 0; w < kEqWarps; ++w) matvec2(bc + kEqcMw, g1, g2, warpAgg[w][0], warpAgg[w][1]);
                        double s1 = 0.0, s2 = 0.0;
                        if (run > 0)
                        {
                            const double* rec = a.chain.rec + ((size_t) ((size_t) seq * a.nRuns + (run - 1)) * CPQ_NUM_BANDS + b) * 4;
                            const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(rec + 2);
                            while (ld_acquire_u64(flag) != a.chain.epoch) { __nanosleep(20); }
                            s1 = ld_cg_f64(rec);
                            s2 = ld_cg_f64(rec + 1);
                        }
                        if (run + 1 < a.nRuns)
                        {
                            double o1 = s1, o2 = s2;
                            matvec2(bc + kEqcMt, o1, o2, g1, g2);
                            double* rec = a.chain.rec + ((size_t) ((size_t) seq * a.nRuns + run) * CPQ_NUM_BANDS + b) * 4;
                            rec[0] = o1;
                            rec[1] = o2;
                            st_release_u64(reinterpret_cast<unsigned long long*>(rec + 2), a.chain.epoch);
                        }
                        sIn[0] = s1; sIn[1] = s2;
                    }
                    __syncthreads();
                    p1 = sIn[0]; p2 = sIn[1];
                }
                else
                {
                    p1 = carry[tpar][b][0]; p2 = carry[tpar][b][1];
                }

                // ---- state before this warp, then before this thread: A^(16 lane) p + e ----
                for (int w = 0; w < warp; ++w) matvec2(bc + kEqcMw, p1, p2, warpAgg[w][0], warpAgg[w][1]);
#pragma unroll
                for (int d = 0; d < 5; ++d)
                {
                    double n1 = p1, n2 = p2;
                    matvec2(bc + kEqcMs + 4 * d, n1, n2, 0.0, 0.0);
                    if ((lane >> d) & 1) { p1 = n1; p2 = n2; }
                }
                double ic1 = p1 + e1, ic2 = p2 + e2;
                const double s1_0 = ic1, s2_0 = ic2;

                // final state of the sequence = state at sample T (T is a multiple of 16)
                if (t0 + kEqTile >= a.T && a.stateOut)
                {
                    const int64_t rem = a.T - t0;   // 1..4096
                    if (rem < kEqTile && tid == (int) (rem / kEqL))
                    {
                        a.stateOut[((size_t) seq * CPQ_NUM_BANDS + b) * 2] = ic1;
                        a.stateOut[((size_t) seq * CPQ_NUM_BANDS + b) * 2 + 1] = ic2;
                    }
                }

                // ---- pass 2, fast path: the reference recurrence (processBandStereo association) ----
                const double a1 = bc[0], a2 = bc[1], a3 = bc[2], m0 = bc[3], m1 = bc[4], m2 = bc[5];
                hiMax = max((unsigned) __double2hiint(ic1) & 0x7fffffffu, (unsigned) __double2hiint(ic2) & 0x7fffffffu);
                bool rare = suspicious | (hiMax >= 0x426d1a94u) | (bc[6] != 0.0);   // |state| >= 1e12
                hiMax = 0;
                const bool peak = bc[7] != 0.0;   // m0 == 1 && m2 == 0
                if (sat > 0.0)
                {
                    if (peak) eq_pass2<true, true>(x, ic1, ic2, a1, a2, a3, m0, m1, m2, alpha, gamma, hiMax);
                    else eq_pass2<true, false>(x, ic1, ic2, a1, a2, a3, m0, m1, m2, alpha, gamma, hiMax);
                }
                else
                {
                    if (peak) eq_pass2<false, true>(x, ic1, ic2, a1, a2, a3, m0, m1, m2, alpha, gamma, hiMax);
                    else eq_pass2<false, false>(x, ic1, ic2, a1, a2, a3, m0, m1, m2, alpha, gamma, hiMax);
                }
                rare |= hiMax >= thrHi;   // some |out| >= 4.5 (100 without saturation), or NaN
                suspicious = false;       // outputs of a band are bounded by 100 (or replayed exactly below)
                if (rare)
                {
                    // exact replay from the stashed inputs and the same start state
                    ic1 = s1_0;
                    ic2 = s2_0;
                    if (eq_exact_block(myStash, ic1, ic2, a1, a2, a3, m0, m1, m2, sat)) atomicExch(a.fault, 1u);
#pragma unroll
                    for (int j = 0; j < kEqL; ++j) x[j] = myStash[j];
                }
                if (tid == kEqThreads - 1)
                {
                    // state after the tile's last sample: carry into the next tile of this run / final state
                    carry[tpar ^ 1][b][0] = ic1;
                    carry[tpar ^ 1][b][1] = ic2;
                    if (a.stateOut && (a.T - t0) == kEqTile)
                    {
                        a.stateOut[((size_t) seq * CPQ_NUM_BANDS + b) * 2] = ic1;
                        a.stateOut[((size_t) seq * CPQ_NUM_BANDS + b) * 2 + 1] = ic2;
                    }
                }
            }
        }

        // ---- store: total gain ramp, makeup, headroom ----
#pragma unroll
        for (int j = 0; j < kEqL; ++j) myStash[j] = x[j];
        __syncthreads();
#pragma unroll 4
        for (int i = tid; i < nValid; i += kEqThreads)
        {
            double v = tile[eq_sidx(i)];
            const int64_t t = t0 + i;
            if (a.doEq)
            {
                if (a.gainTab)
                {
                    const int64_t c = t >> a.blockLog2;
                    const int off = (int) t & bmask;
                    const double2 g = __ldg(reinterpret_cast<const double2*>(a.gainTab) + (size_t) set * a.nCallbacks + c);
                    v *= fma((double) off, g.y, g.x);
                }
                else v *= __ldg(a.gainConst + set);
            }
            if (a.doEpilogue)
            {
                v *= a.makeup;
                if (a.applyHeadroom) v *= 0.8912509381337456;
            }
            io[t] = v;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Dither + 12-tap error-feedback noise shaper, one thread per sequence (serial in time by nature).
// ---------------------------------------------------------------------------------------------
struct DitherArgs
{
    double* io;
    int64_t ioStride;
    int64_t T;
    int nSeq;
    const double* uniforms;   // [nSeq][2*T]
    double coeff[12];
    double scale, invScale;
    double* z;                // [nSeq][12] error history (carried)
};

__global__ void dither_kernel(DitherArgs a)
{
    const int seq = blockIdx.x * blockDim.x + threadIdx.x;
    if (seq >= a.nSeq) return;
    double* d = a.io + (size_t) seq * a.ioStride;
    const double* u = a.uniforms + (size_t) seq * 2 * a.T;
    double z[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) z[i] = a.z[(size_t) seq * 12 + i];
    for (int64_t i = 0; i < a.T; ++i)
    {
        double shaped = a.coeff[0] * z[0];
#pragma unroll
        for (int t = 1; t < 12; ++t) shaped = __dadd_rn(shaped, __dmul_rn(a.coeff[t], z[t]));
        const double2 uu = __ldg(reinterpret_cast<const double2*>(u) + i);
        const double dn = __dmul_rn(__dadd_rn(uu.x - 0.5, uu.y - 0.5), a.scale);
        const double tmp = __dadd_rn(__dadd_rn(__dmul_rn(d[i], 0.8912509381337456), dn), shaped);
        const double q = __dmul_rn(rint(__dmul_rn(tmp, a.invScale)), a.scale);
        double err = __dadd_rn(tmp, -q);
        if (fabs(err) < 1.0e-20) err = 0.0;
#pragma unroll
        for (int t = 11; t > 0; --t) z[t] = z[t - 1];
        z[0] = err;
        d[i] = q;
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) a.z[(size_t) seq * 12 + i] = z[i];
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
