// cpq_kernels.cuh -- hand-written sm_100a kernels of the hot path (FP64 throughout).
//
//   fft_fwd_kernel   batched real->CCS FFT of overlap-save frames (and of IR partitions at prepare)
//                    = ProductionFft::forwardRealToCCS (FFTBackend.cpp:123-135) for every frame at once
//   mac_kernel       Y[k][m] = sum_q X[k-q][m] * H[q][m]: the FDL multiply-accumulate of
//                    processLayerBlock / Add (MKLNonUniformConvolver.cpp:1293-1308, 1505-1520) written as a
//                    Q-tap complex FIR along the frame index, independently per bin
//   fft_inv_kernel   batched CCS->real FFT (1/N), keeps samples [P, 2P) = inverseCCSToR + ringWrite /
//                    tailOutputBuf copy (MKLNonUniformConvolver.cpp:1327-1332, 1531-1540)
//   eq_kernel        layer assembly (Get, :1553-1634) -> 20 x TPT-SVF band with saturation
//                    (EQProcessor.Processing.cpp:191-276) as a blocked linear-recurrence scan -> total gain
//                    ramp (:1262-1274) -> makeup gain + headroom (DSPCoreDouble.cpp:465-469,:655-663)
//   dither_kernel    PsychoacousticDither::processStereoBlock recurrence (PsychoacousticDither.h:293-355)
//
// No tensor cores: none of these stages is a dense contraction.  All HBM access is coalesced double /
// double2; shared memory stages the FFT passes and the EQ tile transposition.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace cpq
{

// ---------------------------------------------------------------------------------------------
// small complex helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
// multiply by -i (SIGN = -1, forward) or +i (SIGN = +1, inverse)
template <int SIGN>
__device__ __forceinline__ double2 mul_i(double2 a)
{
    return SIGN < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
}

// W_P^idx (forward sign) from the layer table tw[t] = exp(-2 pi i t / (2P)), t = 0..P.
template <int SIGN>
__device__ __forceinline__ double2 twiddleP(const double2* __restrict__ tw, int P, int idx)
{
    const int t = 2 * idx;
    double2 w;
    if (t <= P) w = __ldg(tw + t);
    else
    {
        w = __ldg(tw + (t - P));
        w.x = -w.x;
        w.y = -w.y;
    }
    if (SIGN > 0) w.y = -w.y;
    return w;
}

// ---------------------------------------------------------------------------------------------
// radix-2/4/8 DFTs on registers, natural-order output
// ---------------------------------------------------------------------------------------------
template <int SIGN>
__device__ __forceinline__ void dft2(double2& a, double2& b)
{
    const double2 t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

template <int SIGN>
__device__ __forceinline__ void dft4(double2* v)
{
    const double2 b0 = cadd(v[0], v[2]), b2 = csub(v[0], v[2]);
    const double2 b1 = cadd(v[1], v[3]), b3 = mul_i<SIGN>(csub(v[1], v[3]));
    v[0] = cadd(b0, b1);
    v[2] = csub(b0, b1);
    v[1] = cadd(b2, b3);
    v[3] = csub(b2, b3);
}

template <int SIGN>
__device__ __forceinline__ void dft8(double2* v)
{
    constexpr double kS = 0.70710678118654752440;
    const double2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
    const double2 a1 = cadd(v[1], v[5]);
    double2 a5 = csub(v[1], v[5]);
    const double2 a2 = cadd(v[2], v[6]);
    const double2 a6 = mul_i<SIGN>(csub(v[2], v[6]));
    const double2 a3 = cadd(v[3], v[7]);
    double2 a7 = csub(v[3], v[7]);
    // W8^1 = (1 -+ i)/sqrt2, W8^3 = (-1 -+ i)/sqrt2   (upper sign: forward)
    if (SIGN < 0)
    {
        a5 = make_double2((a5.x + a5.y) * kS, (a5.y - a5.x) * kS);
        a7 = make_double2((a7.y - a7.x) * kS, -(a7.x + a7.y) * kS);
    }
    else
    {
        a5 = make_double2((a5.x - a5.y) * kS, (a5.x + a5.y) * kS);
        a7 = make_double2(-(a7.x + a7.y) * kS, (a7.x - a7.y) * kS);
    }
    double2 e[4] = { a0, a1, a2, a3 };
    double2 o[4] = { a4, a5, a6, a7 };
    dft4<SIGN>(e);
    dft4<SIGN>(o);
    v[0] = e[0]; v[2] = e[1]; v[4] = e[2]; v[6] = e[3];
    v[1] = o[0]; v[3] = o[1]; v[5] = o[2]; v[7] = o[3];
}

template <int R, int SIGN>
__device__ __forceinline__ void dftR(double2* v)
{
    if (R == 2) dft2<SIGN>(v[0], v[1]);
    else if (R == 4) dft4<SIGN>(v);
    else dft8<SIGN>(v);
}

// One work item of a Stockham pass: n-point transform, sub-transform length Ns before the pass.
// Reads in[j + r*n/R], twiddles by W_{Ns*R}^{k*r}, R-point DFT, result r goes to (j-k)*R + k + r*Ns.
template <int R, int SIGN, class LoadF>
__device__ __forceinline__ void stockham_load(double2* v, int j, int n, int Ns, const double2* __restrict__ tw, int P, LoadF ld)
{
    const int k = j & (Ns - 1);
    const int stride = n / R;
    const int tscale = n / (Ns * R);
#pragma unroll
    for (int r = 0; r < R; ++r)
    {
        double2 x = ld(j + r * stride);
        if (r > 0 && Ns > 1) x = cmul(x, twiddleP<SIGN>(tw, P, k * r * tscale));
        v[r] = x;
    }
    dftR<R, SIGN>(v);
}

template <int R>
__device__ __forceinline__ int stockham_out_base(int j, int Ns)
{
    const int k = j & (Ns - 1);
    return (j - k) * R + k;
}

// ---------------------------------------------------------------------------------------------
// Forward: frames -> CCS spectra.  One FFT = P/8 threads, each owning 8 complex points.
//   LOG2P = log2 of the complex transform length P (frame length N = 2P real samples).
//   Radix plan: one radix-2^(LOG2P%3) pass first when LOG2P%3 != 0, then radix-8 passes.
// ---------------------------------------------------------------------------------------------
struct FwdArgs
{
    const double* src;      // [nSeq][srcStride] real samples
    int64_t srcStride;
    int64_t frameStart0;    // sample index (within a sequence) where frame 0 starts (may be negative)
    int64_t lo, hi;         // valid sample range [lo, hi); outside -> 0
    int halfOnly;           // 1: only the first P samples of a frame are taken (IR partitions)
    int framesPerSeq;       // K
    int64_t totalFrames;    // nSeq * K
    double2* out;           // [nSeq][outFramesPerSeq][P+1]
    int outFramesPerSeq;    // >= K (row pitch of out per sequence, in frames)
    int outFrameOffset;     // frame f is stored at row f + outFrameOffset
    const double2* tw;      // [P+1]
    double scale;           // applied when applyScale
    int applyScale;
    const double* gain;     // nullable [P+1]
    const double* tilt;     // nullable [P+1]
};

template <int LOG2P>
struct FftCfg
{
    static constexpr int P = 1 << LOG2P;
    static constexpr int TPF = P / 8;                                // threads per FFT
    static constexpr int THREADS = TPF >= 256 ? TPF : 256;
    static constexpr int FPC = THREADS / TPF;                        // frames per CTA
    static constexpr int R0 = 1 << (LOG2P % 3);                      // first-pass radix (1 = none)
    static constexpr int NPASS8 = LOG2P / 3;
    static constexpr size_t SMEM = (size_t) FPC * P * sizeof(double2);
};

template <int LOG2P>
__global__ void __launch_bounds__(FftCfg<LOG2P>::THREADS) fft_fwd_kernel(FwdArgs a)
{
    using C = FftCfg<LOG2P>;
    constexpr int P = C::P, TPF = C::TPF;
    extern __shared__ double2 smem_fft[];
    const int fl = threadIdx.x / TPF;       // local frame
    const int t = threadIdx.x % TPF;
    const int64_t gf = (int64_t) blockIdx.x * C::FPC + fl;
    const bool live = gf < a.totalFrames;
    double2* buf = smem_fft + (size_t) fl * P;
    const int64_t seq = live ? gf / a.framesPerSeq : 0;
    const int f = live ? (int) (gf % a.framesPerSeq) : 0;
    const double* src = a.src + seq * a.srcStride;
    const int64_t base = a.frameStart0 + (int64_t) f * P;
    const bool vec_ok = (a.halfOnly == 0);

    auto gload = [&](int idx) -> double2 {
        // z[idx] = x[2 idx] + i x[2 idx + 1]
        const int64_t g = base + 2 * (int64_t) idx;
        if (!live) return make_double2(0.0, 0.0);
        if (a.halfOnly && 2 * idx >= P) return make_double2(0.0, 0.0);
        if (vec_ok && g >= a.lo && g + 1 < a.hi) return __ldg(reinterpret_cast<const double2*>(src + g));
        double2 z;
        z.x = (g >= a.lo && g < a.hi) ? __ldg(src + g) : 0.0;
        z.y = (g + 1 >= a.lo && g + 1 < a.hi) ? __ldg(src + g + 1) : 0.0;
        return z;
    };
    auto sload = [&](int idx) -> double2 { return buf[idx]; };

    int Ns = 1;
    // ---- first pass: from global ----
    if constexpr (C::R0 > 1)
    {
        constexpr int R = C::R0;
        constexpr int ITEMS = 8 / R;
        double2 v[ITEMS][R];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) stockham_load<R, -1>(v[i], t + i * TPF, P, 1, a.tw, P, gload);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
        {
            const int ob = stockham_out_base<R>(t + i * TPF, 1);
#pragma unroll
            for (int r = 0; r < R; ++r) buf[ob + r] = v[i][r];
        }
        Ns = R;
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < C::NPASS8; ++p)
    {
        double2 v[8];
        if (p == 0 && C::R0 == 1) stockham_load<8, -1>(v, t, P, Ns, a.tw, P, gload);
        else
        {
            stockham_load<8, -1>(v, t, P, Ns, a.tw, P, sload);
            __syncthreads();
        }
        const int ob = stockham_out_base<8>(t, Ns);
#pragma unroll
        for (int r = 0; r < 8; ++r) buf[ob + r * Ns] = v[r];
        Ns *= 8;
        __syncthreads();
    }

    // ---- split post-process: X[m] = E + W_N^m O, X[P-m] = conj(E - W_N^m O) ----
    if (!live) return;
    double2* out = a.out + ((size_t) seq * a.outFramesPerSeq + (size_t) (f + a.outFrameOffset)) * (size_t) (P + 1);
    auto emit = [&](int m, double2 X) {
        if (a.applyScale) { X.x *= a.scale; X.y *= a.scale; }
        if (a.gain) { const double g = __ldg(a.gain + m); X.x *= g; X.y *= g; }
        if (a.tilt) { const double g = __ldg(a.tilt + m); X.x *= g; X.y *= g; }
        out[m] = X;
    };
    for (int m = t; m <= P / 2; m += TPF)
    {
        const double2 zm = buf[m];
        const double2 zc = cconj(buf[(P - m) & (P - 1)]);
        const double2 E = make_double2(0.5 * (zm.x + zc.x), 0.5 * (zm.y + zc.y));
        const double2 D = make_double2(0.5 * (zm.x - zc.x), 0.5 * (zm.y - zc.y));
        const double2 O = make_double2(D.y, -D.x);   // -i * D
        const double2 w = __ldg(a.tw + m);
        const double2 Tm = cmul(w, O);
        if (m == 0)
        {
            emit(0, make_double2(E.x + O.x, 0.0));
            emit(P, make_double2(E.x - O.x, 0.0));
        }
        else
        {
            emit(m, cadd(E, Tm));
            if (m != P / 2) emit(P - m, cconj(csub(E, Tm)));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Inverse: CCS spectra -> real, 1/N, keep [P, 2P).
// ---------------------------------------------------------------------------------------------
struct InvArgs
{
    const double2* in;      // [nSeq][framesPerSeq][P+1]
    int framesPerSeq;       // K (row pitch of `in`)
    int framesOut;          // frames f < framesOut are transformed per sequence
    int64_t totalFrames;    // nSeq * framesOut
    double* out;            // [nSeq][outStride]; frame f -> out[f*P .. (f+1)*P)
    int64_t outStride;
    const double2* tw;
};

template <int LOG2P>
__global__ void __launch_bounds__(FftCfg<LOG2P>::THREADS) fft_inv_kernel(InvArgs a)
{
    using C = FftCfg<LOG2P>;
    constexpr int P = C::P, TPF = C::TPF;
    extern __shared__ double2 smem_fft[];
    const int fl = threadIdx.x / TPF;
    const int t = threadIdx.x % TPF;
    const int64_t gf = (int64_t) blockIdx.x * C::FPC + fl;
    const bool live = gf < a.totalFrames;
    double2* buf = smem_fft + (size_t) fl * P;
    const int64_t seq = live ? gf / a.framesOut : 0;
    const int f = live ? (int) (gf % a.framesOut) : 0;
    const double2* Y = a.in + ((size_t) seq * a.framesPerSeq + (size_t) f) * (size_t) (P + 1);
    const double invN = 1.0 / (double) (2 * P);

    // Z[m] = ((Y[m] + conj Y[P-m]) + i conj(W_N^m) (Y[m] - conj Y[P-m])) / N
    auto gload = [&](int m) -> double2 {
        if (!live) return make_double2(0.0, 0.0);
        double2 ym = __ldg(Y + m);
        double2 yc = cconj(__ldg(Y + (P - m)));
        if (m == 0) { ym.y = 0.0; yc.y = 0.0; }   // imaginary parts of bins 0 and P are ignored (CCS contract)
        const double2 S = cadd(ym, yc);
        const double2 D = csub(ym, yc);
        const double2 wc = cconj(__ldg(a.tw + m));
        const double2 Tm = cmul(wc, D);
        // + i * Tm
        return make_double2((S.x - Tm.y) * invN, (S.y + Tm.x) * invN);
    };
    auto sload = [&](int idx) -> double2 { return buf[idx]; };

    int Ns = 1;
    if constexpr (C::R0 > 1)
    {
        constexpr int R = C::R0;
        constexpr int ITEMS = 8 / R;
        double2 v[ITEMS][R];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) stockham_load<R, +1>(v[i], t + i * TPF, P, 1, a.tw, P, gload);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
        {
            const int ob = stockham_out_base<R>(t + i * TPF, 1);
#pragma unroll
            for (int r = 0; r < R; ++r) buf[ob + r] = v[i][r];
        }
        Ns = R;
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < C::NPASS8; ++p)
    {
        double2 v[8];
        if (p == 0 && C::R0 == 1) stockham_load<8, +1>(v, t, P, Ns, a.tw, P, gload);
        else
        {
            stockham_load<8, +1>(v, t, P, Ns, a.tw, P, sload);
            if (p != C::NPASS8 - 1) __syncthreads();
        }
        if (p == C::NPASS8 - 1)
        {
            // last pass: Ns == P/8, outputs at t + r*P/8; only z[P/2 ..) = y[P .. 2P) is kept.
            if (live)
            {
                double2* o = reinterpret_cast<double2*>(a.out + seq * a.outStride + (int64_t) f * P);
#pragma unroll
                for (int r = 4; r < 8; ++r) o[t + (r - 4) * (P / 8)] = v[r];
            }
        }
        else
        {
            const int ob = stockham_out_base<8>(t, Ns);
#pragma unroll
            for (int r = 0; r < 8; ++r) buf[ob + r * Ns] = v[r];
            Ns *= 8;
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Spectrum multiply-accumulate: Y[k][m] = sum_{q in [qBegin,qEnd)} X[k-q][m] * H[q][m]
// Register-blocked KT outputs x QT taps per thread; one thread per bin; X rows with k-q < 0 are the
// zero history before t = 0 (Reset state).
// ---------------------------------------------------------------------------------------------
struct MacArgs
{
    const double2* X;    // [nSeq][K][M]
    const double2* H;    // [nH][Q][M]   natural partition order q = 0..Q-1
    double2* Y;          // [nSeq][K][M]
    int K, M, Q;
    int qBegin, qEnd;
    int64_t hSeqStride;  // elements between sequences in H (0 when every stream shares the IR pair)
    int hSeqMod;         // H row = (seq % hSeqMod) when shared per channel; 0 = seq
};

template <int KT, int QT>
__global__ void __launch_bounds__(128) mac_kernel(MacArgs a)
{
    const int m = blockIdx.x * 128 + threadIdx.x;
    const int k0 = blockIdx.y * KT;
    const int seq = blockIdx.z;
    if (m >= a.M) return;
    const int hrow = a.hSeqMod > 0 ? (seq % a.hSeqMod) : seq;
    const double2* __restrict__ X = a.X + (size_t) seq * a.K * a.M + m;
    const double2* __restrict__ H = a.H + (size_t) hrow * a.hSeqStride + m;
    double2 acc[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) acc[i] = make_double2(0.0, 0.0);

    for (int q0 = a.qBegin; q0 < a.qEnd; q0 += QT)
    {
        double2 h[QT];
#pragma unroll
        for (int i = 0; i < QT; ++i)
            h[i] = (q0 + i < a.qEnd) ? __ldg(H + (size_t) (q0 + i) * a.M) : make_double2(0.0, 0.0);
        const int fLo = k0 - (q0 + QT - 1);
        double2 xw[KT + QT - 1];
#pragma unroll
        for (int i = 0; i < KT + QT - 1; ++i)
        {
            const int f = fLo + i;
            xw[i] = (f >= 0 && f < a.K) ? __ldg(X + (size_t) f * a.M) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int kk = 0; kk < KT; ++kk)
#pragma unroll
            for (int i = 0; i < QT; ++i)
            {
                const double2 x = xw[kk - i + QT - 1];
                acc[kk].x = fma(x.x, h[i].x, acc[kk].x);
                acc[kk].x = fma(-x.y, h[i].y, acc[kk].x);
                acc[kk].y = fma(x.x, h[i].y, acc[kk].y);
                acc[kk].y = fma(x.y, h[i].x, acc[kk].y);
            }
    }
    double2* __restrict__ Y = a.Y + (size_t) seq * a.K * a.M + m;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
        if (k0 + kk < a.K) Y[(size_t) (k0 + kk) * a.M] = acc[kk];
}

// ---------------------------------------------------------------------------------------------
// EQ: assembly -> 20 bands -> gain ramp -> makeup/headroom.
// ---------------------------------------------------------------------------------------------
constexpr int kEqThreads = 256;
constexpr int kEqL = 16;                       // samples per thread
constexpr int kEqTile = kEqThreads * kEqL;     // 4096
constexpr int kEqWarps = kEqThreads / 32;

// per (parameter set, band) constants, all double; see EngineImpl::buildEqConstants for the layout
constexpr int kEqcCoef = 0;      // a1,a2,a3,m0,m1,m2
constexpr int kEqcW = 8;         // w[16][2]   zero-state weights, c = sum_j w[j] * v0[j]
constexpr int kEqcTl = 40;       // Tl[32][4]  A^(16*lane), row-major 2x2
constexpr int kEqcMs = 168;      // Ms[5][4]   A^(16*2^d)
constexpr int kEqcMw = 188;      // A^512
constexpr int kEqcMt = 192;      // A^4096
constexpr int kEqcStride = 196;  // doubles per band

struct EqChain
{
    // flags[(seq*nRuns + run)*20 + band] : {double s1, s2; uint64 epoch}
    double* rec;          // 4 doubles per record (s1, s2, epoch-as-u64, pad)
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
    int assemble;           // add tails
    int nTail;              // number of tail layers (0..2)
    const double* tail[2];  // [nSeq][tailStride[l]] layer output streams
    int64_t tailStride[2];
    const int64_t* tailSrc[2];     // [nCallbacks] stream position or -1
    const int32_t* blockMap[2];    // nullable: stream block -> frame
    int tailPart[2];
    double tailGain[2];
    int blockSize;
    int outer;              // CPQ_CONV_OUTER: scrub + wet gain
    double wetGain;
    // EQ
    int doEq;
    const double* eqc;      // [nSets][20][kEqcStride]
    const unsigned* bandMask;  // [nSeq] bit b = band b processed for this sequence
    int eqSetMod;           // set = shared ? 0 : seq / channels ... resolved by setOfSeq
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
    unsigned* fault;        // set to 1 when a state left the linear regime (|ic| >= 1e15 or non-finite)
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

// 1/x to ~1 ulp: hardware reciprocal seed + two Newton steps (FP64 pipe), no slow-path branch.
__device__ __forceinline__ double fast_div(double num, double den)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    double e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    double q = num * r;
    const double rem = fma(-den, q, num);
    q = fma(rem, r, q);
    return q;
}

__device__ __forceinline__ bool eq_valid(double v) { return fabs(v) < 1.0e15; }   // false for NaN / Inf too

// padded shared index: 17-double stride per 16 samples keeps both the coalesced pass (consecutive t) and
// the per-thread pass (16 consecutive samples per lane) free of bank conflicts
__device__ __forceinline__ int eq_sidx(int t) { return t + (t >> 4); }

__global__ void __launch_bounds__(kEqThreads) eq_kernel(EqArgs a)
{
    __shared__ double tile[kEqTile + kEqTile / 16];
    __shared__ double warpAggBuf[2][kEqWarps][2];   // double-buffered by band parity
    __shared__ double sIn[2];
    __shared__ double carry[CPQ_NUM_BANDS][2];
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
    const double oneMinusSat = 1.0 - sat;
    const double* eqcSet = a.eqc + (size_t) set * CPQ_NUM_BANDS * kEqcStride;
    const bool chained = a.tilesPerRun == 1 && a.nRuns > 1;

    if (tid < CPQ_NUM_BANDS * 2) (&carry[0][0])[tid] = 0.0;
    __syncthreads();

    for (int tl = 0; tl < a.tilesPerRun; ++tl)
    {
        const int tileIdx = run * a.tilesPerRun + tl;
        if (tileIdx >= a.nTiles) break;
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
                    const int64_t c = t / a.blockSize;
                    const int off = (int) (t - c * a.blockSize);
                    for (int l = 0; l < a.nTail; ++l)
                    {
                        const int64_t s = __ldg(a.tailSrc[l] + c);
                        if (s >= 0)
                        {
                            int64_t pos = s + off;
                            if (a.blockMap[l])
                            {
                                const int64_t j = pos / a.tailPart[l];
                                pos = (int64_t) __ldg(a.blockMap[l] + j) * a.tailPart[l] + (pos - j * a.tailPart[l]);
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
#pragma unroll
        for (int j = 0; j < kEqL; ++j) x[j] = tile[eq_sidx(tid * kEqL + j)];

        if (a.doEq)
        {
            int parity = 0;
            for (int b = 0; b < CPQ_NUM_BANDS; ++b)
            {
                if (!((mask >> b) & 1u)) continue;   // uniform per CTA
                double (*warpAgg)[2] = warpAggBuf[parity];
                parity ^= 1;
                const double* __restrict__ bc = eqcSet + (size_t) b * kEqcStride;
                // ---- pass 1: zero-state response of this thread's 16 samples ----
                double c1 = 0.0, c2 = 0.0;
#pragma unroll
                for (int j = 0; j < kEqL; ++j)
                {
                    const double2 w = __ldg(reinterpret_cast<const double2*>(bc + kEqcW) + j);
                    c1 = fma(w.x, x[j], c1);
                    c2 = fma(w.y, x[j], c2);
                }
                // ---- warp inclusive scan of s -> A^16 s + c ----
#pragma unroll
                for (int d = 0; d < 5; ++d)
                {
                    const double p1 = __shfl_up_sync(0xffffffffu, c1, 1 << d);
                    const double p2 = __shfl_up_sync(0xffffffffu, c2, 1 << d);
                    if (lane >= (1 << d))
                    {
                        const double2 m01 = __ldg(reinterpret_cast<const double2*>(bc + kEqcMs + 4 * d));
                        const double2 m23 = __ldg(reinterpret_cast<const double2*>(bc + kEqcMs + 4 * d) + 1);
                        c1 = fma(m01.x, p1, fma(m01.y, p2, c1));
                        c2 = fma(m23.x, p1, fma(m23.y, p2, c2));
                    }
                }
                if (lane == 31) { warpAgg[warp][0] = c1; warpAgg[warp][1] = c2; }
                // exclusive value (state contribution before this thread, relative to the warp start)
                double e1 = __shfl_up_sync(0xffffffffu, c1, 1);
                double e2 = __shfl_up_sync(0xffffffffu, c2, 1);
                if (lane == 0) { e1 = 0.0; e2 = 0.0; }
                __syncthreads();

                const double2 mw01 = __ldg(reinterpret_cast<const double2*>(bc + kEqcMw));
                const double2 mw23 = __ldg(reinterpret_cast<const double2*>(bc + kEqcMw) + 1);
                if (tid == 0)
                {
                    // tile aggregate with zero carry-in, then the carry-in itself
                    double g1 = 0.0, g2 = 0.0;
                    for (int w = 0; w < kEqWarps; ++w)
                    {
                        const double n1 = fma(mw01.x, g1, fma(mw01.y, g2, warpAgg[w][0]));
                        const double n2 = fma(mw23.x, g1, fma(mw23.y, g2, warpAgg[w][1]));
                        g1 = n1; g2 = n2;
                    }
                    double s1 = carry[b][0], s2 = carry[b][1];
                    if (chained && tl == 0 && run > 0)
                    {
                        const double* rec = a.chain.rec + ((size_t) ((size_t) seq * a.nRuns + (run - 1)) * CPQ_NUM_BANDS + b) * 4;
                        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(rec + 2);
                        while (ld_acquire_u64(flag) != a.chain.epoch) { __nanosleep(20); }
                        s1 = ld_cg_f64(rec);
                        s2 = ld_cg_f64(rec + 1);
                    }
                    const double2 mt01 = __ldg(reinterpret_cast<const double2*>(bc + kEqcMt));
                    const double2 mt23 = __ldg(reinterpret_cast<const double2*>(bc + kEqcMt) + 1);
                    const double o1 = fma(mt01.x, s1, fma(mt01.y, s2, g1));
                    const double o2 = fma(mt23.x, s1, fma(mt23.y, s2, g2));
                    if (chained && run + 1 < a.nRuns)
                    {
                        double* rec = a.chain.rec + ((size_t) ((size_t) seq * a.nRuns + run) * CPQ_NUM_BANDS + b) * 4;
                        rec[0] = o1;
                        rec[1] = o2;
                        st_release_u64(reinterpret_cast<unsigned long long*>(rec + 2), a.chain.epoch);
                    }
                    sIn[0] = s1; sIn[1] = s2;
                    carry[b][0] = o1; carry[b][1] = o2;   // carry into the next tile of this run
                }
                __syncthreads();

                // ---- state before this warp, then before this thread ----
                double p1 = sIn[0], p2 = sIn[1];
                for (int w = 0; w < warp; ++w)
                {
                    const double n1 = fma(mw01.x, p1, fma(mw01.y, p2, warpAgg[w][0]));
                    const double n2 = fma(mw23.x, p1, fma(mw23.y, p2, warpAgg[w][1]));
                    p1 = n1; p2 = n2;
                }
                const double2 tl01 = __ldg(reinterpret_cast<const double2*>(bc + kEqcTl + 4 * lane));
                const double2 tl23 = __ldg(reinterpret_cast<const double2*>(bc + kEqcTl + 4 * lane) + 1);
                double ic1 = fma(tl01.x, p1, fma(tl01.y, p2, e1));
                double ic2 = fma(tl23.x, p1, fma(tl23.y, p2, e2));

                // final state of the sequence = state at sample T
                if (t0 + kEqTile >= a.T && a.stateOut)
                {
                    const int64_t rem = a.T - t0;   // 1..4096, multiple of 16
                    if (rem < kEqTile && tid == (int) (rem / kEqL))
                    {
                        a.stateOut[((size_t) seq * CPQ_NUM_BANDS + b) * 2] = ic1;
                        a.stateOut[((size_t) seq * CPQ_NUM_BANDS + b) * 2 + 1] = ic2;
                    }
                }

                // ---- pass 2: the reference recurrence (processBandStereo association) ----
                const double a1 = __ldg(bc + 0), a2 = __ldg(bc + 1), a3 = __ldg(bc + 2);
                const double m0 = __ldg(bc + 3), m1 = __ldg(bc + 4), m2 = __ldg(bc + 5);
                bool bad = false;
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
                        double xc = fmax(out, -4.5);   // NaN -> -4.5 like _mm_max_pd(x, lo)
                        xc = fmin(xc, 4.5);
                        const double x2 = xc * xc;
                        const double th = fast_div(xc * (27.0 + x2), fma(9.0, x2, 27.0));
                        out = out * oneMinusSat + th * sat;
                    }
                    if (!eq_valid(out)) out = 0.0;
                    bad |= !eq_valid(ic1) | !eq_valid(ic2);
                    out = fmin(fmax(out, -100.0), 100.0);
                    x[j] = out;
                }
                if (bad) atomicExch(a.fault, 1u);
                if (t0 + kEqTile >= a.T && a.stateOut && (a.T - t0) == kEqTile && tid == kEqThreads - 1)
                {
                    a.stateOut[((size_t) seq * CPQ_NUM_BANDS + b) * 2] = ic1;
                    a.stateOut[((size_t) seq * CPQ_NUM_BANDS + b) * 2 + 1] = ic2;
                }
            }
        }

        // ---- store: total gain ramp, makeup, headroom ----
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kEqL; ++j) tile[eq_sidx(tid * kEqL + j)] = x[j];
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
                    const int64_t c = t / a.blockSize;
                    const int off = (int) (t - c * a.blockSize);
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

} // namespace cpq
