// cpq_fft.cuh -- batched FP64 real FFTs for the overlap-save frames (sm_100a).
//
//   fft_fwd_kernel   real -> CCS spectra of overlap-save frames (and of IR partitions at prepare)
//                    = ProductionFft::forwardRealToCCS (FFTBackend.cpp:123-135) for every frame at once
//   fft_inv_kernel   CCS -> real (1/N), keeps samples [P, 2P) = inverseCCSToR + ringWrite / tailOutputBuf
//                    copy (MKLNonUniformConvolver.cpp:1327-1332, 1531-1540)
//
// A 2P-point real transform is a P-point complex transform of z[n] = x[2n] + i x[2n+1] plus a split
// pass.  The complex transform is a Stockham autosort FFT: one radix-2/4 pass when log2 P is not a
// multiple of 3, then radix-8 passes; each thread owns 8 points in registers per pass, passes exchange
// through shared memory.  The first pass reads HBM directly (coalesced double2), the forward split pass and
// the inverse's last pass write HBM directly, so every sample crosses HBM once in each direction.
// Shared memory is indexed through pad(i) = i + (i >> 3): one 16-byte slot of padding per 8 keeps the
// stride-8 writes of the first pass, the stride-1 accesses of later passes and the mirrored reads of the
// split pass free of bank conflicts (ncu round 1: 51 % of shared wavefronts were conflicts without it).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace cpq
{

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

__device__ __forceinline__ int fft_pad(int i) { return i + (i >> 3); }

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
// Radix-8 passes take W^k, W^2k, W^4k from a compact per-pass table (unit stride in k, so the loads of a warp
// coalesce) and form the other four powers by multiplication; ptw = {T1[Ns], T2[Ns], T4[Ns]} for this pass.
template <int R, int SIGN, class LoadF>
__device__ __forceinline__ void stockham_load(double2* v, int j, int n, int Ns, const double2* __restrict__ ptw, LoadF ld)
{
    const int k = j & (Ns - 1);
    const int stride = n / R;
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = ld(j + r * stride);
    if (R == 8 && Ns > 1)
    {
        double2 w1 = __ldg(ptw + k), w2 = __ldg(ptw + Ns + k), w4 = __ldg(ptw + 2 * Ns + k);
        if (SIGN > 0) { w1.y = -w1.y; w2.y = -w2.y; w4.y = -w4.y; }
        const double2 w3 = cmul(w1, w2), w5 = cmul(w1, w4), w6 = cmul(w2, w4);
        const double2 w7 = cmul(w3, w4);
        v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3); v[4] = cmul(v[4], w4);
        v[5] = cmul(v[5], w5); v[6] = cmul(v[6], w6); v[7] = cmul(v[7], w7);
    }
    dftR<R, SIGN>(v);
}

template <int R>
__device__ __forceinline__ int stockham_out_base(int j, int Ns)
{
    const int k = j & (Ns - 1);
    return (j - k) * R + k;
}

struct FwdArgs
{
    const double* src;      // [nSeq][srcStride] real samples
    int64_t srcStride;
    int64_t frameStart0;    // sample index (within a sequence) where frame 0 starts (may be negative)
    int64_t lo, hi;         // valid sample range [lo, hi); outside -> 0
    int halfOnly;           // 1: only the first P samples of a frame are taken (IR partitions)
    int framesPerSeq;       // K
    int64_t totalFrames;    // nSeq * K
    double2* out;           // [nSeq][outFramesPerSeq][P]  packed: slot 0 = (Re X[0], Re X[P])
    int outFramesPerSeq;    // >= K (row pitch of out per sequence, in frames)
    int outFrameOffset;     // frame f is stored at row f + outFrameOffset
    const double2* tw;      // [P+1]  exp(-2 pi i t / 2P), split pass
    const double2* ptw;     // per-pass radix-8 tables, see FftCfg::passOffset
    double scale;           // applied when applyScale
    int applyScale;
    const double* gain;     // nullable [P+1]
    const double* tilt;     // nullable [P+1]
    double2* scratch;       // only for P > 8192 (cpq_fft_large.cuh): same geometry as out
    // streaming continuation: samples before the call come from the carried input history instead of a concatenated copy.
    // histEnd points one past the last history sample of sequence 0 (so histEnd[seq * histStride + g] is sample g < 0);
    // null = no history, and then lo >= 0
    const double* histEnd;
    int64_t histStride;
};

// sample g of a sequence: this call's row for g >= 0, the carried history for g < 0, zero outside [lo, hi)
__device__ __forceinline__ double fwd_sample(const double* __restrict__ src, const double* __restrict__ hist, int64_t lo, int64_t hi, int64_t g)
{
    if (g < lo || g >= hi) return 0.0;
    return g >= 0 ? __ldg(src + g) : (hist ? __ldg(hist + g) : 0.0);
}

template <int LOG2P>
struct FftCfg
{
    static constexpr int P = 1 << LOG2P;
    static constexpr int TPF = P / 8;                                // threads per FFT
    static constexpr int THREADS = TPF >= 256 ? TPF : 256;
    static constexpr int FPC = THREADS / TPF;                        // frames per CTA
    static constexpr int R0 = 1 << (LOG2P % 3);                      // first-pass radix (1 = none)
    static constexpr int NPASS8 = LOG2P / 3;
    static constexpr int ROW = P + P / 8;                            // padded row, in double2
    static constexpr size_t SMEM = (size_t) FPC * ROW * sizeof(double2);
    static constexpr int MINBLOCKS = THREADS <= 256 ? 4 : (THREADS <= 512 ? 2 : 1);
    // radix-8 pass p works on sub-transform length Ns = R0 * 8^p; its table {T1,T2,T4}[Ns] starts here
    __host__ __device__ static constexpr int passNs(int p) { int ns = R0; for (int i = 0; i < p; ++i) ns *= 8; return ns; }
    __host__ __device__ static constexpr int passOffset(int p)
    {
        int off = 0;
        for (int i = 0; i < p; ++i) if (passNs(i) > 1) off += 3 * passNs(i);
        return off;
    }
    static constexpr int PTW_SIZE = passOffset(NPASS8);
};

// Frames of one CTA never exchange data: each frame's TPF threads (whole warps) meet at their own named barrier.
template <int TPF, int THREADS>
__device__ __forceinline__ void fft_sync(int fl)
{
    if constexpr (TPF == THREADS || (TPF % 32) != 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(fl + 1), "n"(TPF) : "memory");
}

// IR = prepare-time variant (half frames, scale, spectrum gain / tilt); the streaming variant carries none of that.
template <int LOG2P, bool IR>
__global__ void __launch_bounds__(FftCfg<LOG2P>::THREADS, FftCfg<LOG2P>::MINBLOCKS) fft_fwd_kernel(FwdArgs a)
{
    using C = FftCfg<LOG2P>;
    constexpr int P = C::P, TPF = C::TPF;
    extern __shared__ double2 smem_fft[];
    const int fl = threadIdx.x / TPF;       // local frame
    const int t = threadIdx.x % TPF;
    const int64_t gf = (int64_t) blockIdx.x * C::FPC + fl;
    const bool live = gf < a.totalFrames;
    double2* buf = smem_fft + (size_t) fl * C::ROW;
    const int64_t seq = live ? gf / a.framesPerSeq : 0;
    const int f = live ? (int) (gf % a.framesPerSeq) : 0;
    const double* src = a.src + seq * a.srcStride;
    const int64_t base = a.frameStart0 + (int64_t) f * P;
    const bool vec_ok = !IR || (a.halfOnly == 0);
    // whole frame inside the valid range and 16-byte aligned: no per-element checks (uniform per frame)
    const double* hist = a.histEnd ? a.histEnd + seq * a.histStride : nullptr;
    const bool inside = !IR && live && base >= 0 && base >= a.lo && base + 2 * P <= a.hi && ((reinterpret_cast<uintptr_t>(src + base) & 15) == 0);
    const double2* src2 = reinterpret_cast<const double2*>(src + base);

    auto gload = [&](int idx) -> double2 {
        // z[idx] = x[2 idx] + i x[2 idx + 1]
        if (inside) return __ldg(src2 + idx);
        const int64_t g = base + 2 * (int64_t) idx;
        if (!live) return make_double2(0.0, 0.0);
        if (IR && a.halfOnly && 2 * idx >= P) return make_double2(0.0, 0.0);
        if (vec_ok && g >= 0 && g >= a.lo && g + 1 < a.hi && ((reinterpret_cast<uintptr_t>(src + g) & 15) == 0)) return __ldg(reinterpret_cast<const double2*>(src + g));
        double2 z;
        z.x = fwd_sample(src, hist, a.lo, a.hi, g);
        z.y = fwd_sample(src, hist, a.lo, a.hi, g + 1);
        return z;
    };
    auto sload = [&](int idx) -> double2 { return buf[fft_pad(idx)]; };

    int Ns = 1;
    if constexpr (C::R0 > 1)
    {
        constexpr int R = C::R0;
        constexpr int ITEMS = 8 / R;
        double2 v[ITEMS][R];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) stockham_load<R, -1>(v[i], t + i * TPF, P, 1, a.ptw, gload);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
        {
            const int ob = stockham_out_base<R>(t + i * TPF, 1);
#pragma unroll
            for (int r = 0; r < R; ++r) buf[fft_pad(ob + r)] = v[i][r];
        }
        Ns = R;
        fft_sync<TPF, C::THREADS>(fl);
    }
#pragma unroll
    for (int p = 0; p < C::NPASS8; ++p)
    {
        double2 v[8];
        if (p == 0 && C::R0 == 1) stockham_load<8, -1>(v, t, P, Ns, a.ptw + C::passOffset(p), gload);
        else
        {
            stockham_load<8, -1>(v, t, P, Ns, a.ptw + C::passOffset(p), sload);
            fft_sync<TPF, C::THREADS>(fl);
        }
        const int ob = stockham_out_base<8>(t, Ns);
#pragma unroll
        for (int r = 0; r < 8; ++r) buf[fft_pad(ob + r * Ns)] = v[r];
        Ns *= 8;
        fft_sync<TPF, C::THREADS>(fl);
    }

    // ---- split pass: X[m] = E + W_N^m O, X[P-m] = conj(E - W_N^m O) ----
    if (!live) return;
    double2* out = a.out + ((size_t) seq * a.outFramesPerSeq + (size_t) (f + a.outFrameOffset)) * (size_t) P;   // packed row
    auto emit = [&](int m, double2 X) {
        if constexpr (IR)
        {
            if (a.applyScale) { X.x *= a.scale; X.y *= a.scale; }
            if (a.gain) { const double g = __ldg(a.gain + m); X.x *= g; X.y *= g; }
            if (a.tilt) { const double g = __ldg(a.tilt + m); X.x *= g; X.y *= g; }
        }
        out[m] = X;
    };
    // with S = z[m] + conj z[P-m], D = z[m] - conj z[P-m] and the table twF[m] = -i/2 W_N^m:  T = twF[m] D,
    // X[m] = S/2 + T, X[P-m] = conj(S/2 - T)
    const double2* __restrict__ twF = a.tw + (P + 1);
#pragma unroll
    for (int i = 0; i <= (P / 2) / TPF; ++i)
    {
        const int m = t + i * TPF;
        if (m > P / 2) break;
        const double2 zm = buf[fft_pad(m)];
        if (m == 0)
        {
            // bins 0 and P are real: packed into slot 0 as (Re X[0], Re X[P])
            double r0 = zm.x + zm.y, rP = zm.x - zm.y;
            if constexpr (IR)
            {
                if (a.applyScale) { r0 *= a.scale; rP *= a.scale; }
                if (a.gain) { r0 *= __ldg(a.gain); rP *= __ldg(a.gain + P); }
                if (a.tilt) { r0 *= __ldg(a.tilt); rP *= __ldg(a.tilt + P); }
            }
            out[0] = make_double2(r0, rP);
            continue;
        }
        const double2 zp = buf[fft_pad(P - m)];
        const double2 S = make_double2(zm.x + zp.x, zm.y - zp.y);
        const double2 D = make_double2(zm.x - zp.x, zm.y + zp.y);
        const double2 T = cmul(__ldg(twF + m), D);
        emit(m, make_double2(fma(0.5, S.x, T.x), fma(0.5, S.y, T.y)));
        if (m != P / 2) emit(P - m, make_double2(fma(0.5, S.x, -T.x), fma(-0.5, S.y, T.y)));
    }
}

struct InvArgs
{
    const double2* in;      // [nSeq][framesPerSeq][P] packed
    int framesPerSeq;       // K (row pitch of `in`)
    int framesOut;          // frames f < framesOut are transformed per sequence
    int64_t totalFrames;    // nSeq * framesOut
    double* out;            // [nSeq][outStride]; frame f -> out[f*P .. (f+1)*P)
    int64_t outStride;
    const double2* tw;
    const double2* ptw;
    double2* scratch;       // only for P > 8192: same geometry as in (which is overwritten)
};

template <int LOG2P>
__global__ void __launch_bounds__(FftCfg<LOG2P>::THREADS, FftCfg<LOG2P>::MINBLOCKS) fft_inv_kernel(InvArgs a)
{
    using C = FftCfg<LOG2P>;
    constexpr int P = C::P, TPF = C::TPF;
    extern __shared__ double2 smem_fft[];
    const int fl = threadIdx.x / TPF;
    const int t = threadIdx.x % TPF;
    const int64_t gf = (int64_t) blockIdx.x * C::FPC + fl;
    const bool live = gf < a.totalFrames;
    double2* buf = smem_fft + (size_t) fl * C::ROW;
    const int64_t seq = live ? gf / a.framesOut : 0;
    const int f = live ? (int) (gf % a.framesOut) : 0;
    const double2* Y = a.in + ((size_t) seq * a.framesPerSeq + (size_t) f) * (size_t) P;   // packed row
    const double invN = 1.0 / (double) (2 * P);

    // Z[m] = ((Y[m] + conj Y[P-m]) + i conj(W_N^m) (Y[m] - conj Y[P-m])) / N = S/N + twI[m] D,  twI[m] = i conj(W_N^m) / N
    const double2* __restrict__ twI = a.tw + 2 * (P + 1);
    auto gload = [&](int m) -> double2 {
        if (!live) return make_double2(0.0, 0.0);
        double2 ym, yc;
        if (m == 0)
        {
            const double2 y0 = __ldg(Y);   // packed (Re Y[0], Re Y[P]); imaginary parts of bins 0 and P do not exist
            ym = make_double2(y0.x, 0.0);
            yc = make_double2(y0.y, 0.0);
        }
        else
        {
            ym = __ldg(Y + m);
            yc = cconj(__ldg(Y + (P - m)));
        }
        const double2 S = cadd(ym, yc);
        const double2 D = csub(ym, yc);
        const double2 T = cmul(__ldg(twI + m), D);
        return make_double2(fma(S.x, invN, T.x), fma(S.y, invN, T.y));
    };
    auto sload = [&](int idx) -> double2 { return buf[fft_pad(idx)]; };

    int Ns = 1;
    if constexpr (C::R0 > 1)
    {
        constexpr int R = C::R0;
        constexpr int ITEMS = 8 / R;
        double2 v[ITEMS][R];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) stockham_load<R, +1>(v[i], t + i * TPF, P, 1, a.ptw, gload);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
        {
            const int ob = stockham_out_base<R>(t + i * TPF, 1);
#pragma unroll
            for (int r = 0; r < R; ++r) buf[fft_pad(ob + r)] = v[i][r];
        }
        Ns = R;
        fft_sync<TPF, C::THREADS>(fl);
    }
#pragma unroll
    for (int p = 0; p < C::NPASS8; ++p)
    {
        double2 v[8];
        if (p == 0 && C::R0 == 1) stockham_load<8, +1>(v, t, P, Ns, a.ptw + C::passOffset(p), gload);
        else
        {
            stockham_load<8, +1>(v, t, P, Ns, a.ptw + C::passOffset(p), sload);
            if (p != C::NPASS8 - 1) fft_sync<TPF, C::THREADS>(fl);
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
            for (int r = 0; r < 8; ++r) buf[fft_pad(ob + r * Ns)] = v[r];
            Ns *= 8;
            fft_sync<TPF, C::THREADS>(fl);
        }
    }
}

} // namespace cpq
