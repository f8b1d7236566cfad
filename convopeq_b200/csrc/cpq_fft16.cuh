// cpq_fft16.cuh -- the streaming real FFTs for P = 8^n (512, 4096): sixteen points per thread (sm_100a, FP64).
//
// Same transforms as cpq_fft.cuh (ProductionFft::forwardRealToCCS / inverseCCSToR, FFTBackend.cpp:123-151), arranged so
// that the real-FFT split pass (forward) and pre-pass (inverse) need no exchange at all: a thread runs two radix-8
// butterflies per pass, and in the pass that touches the spectrum it takes butterfly j together with butterfly
// P/8 - j.  Butterfly j of that pass produces (consumes) Z[j + r P/8], r = 0..7, whose mirror Z[P - m] is element
// 7 - r of butterfly P/8 - j: every (Z[m], Z[P-m]) pair lives in one thread's registers.  Thread 0 owns the two
// self-mirrored butterflies (0 and P/16) through a register permutation, bins 0 and P/2 being its special pair.
// Compared with eight points per thread this removes one of the shared-memory round trips (the kernels are bound by
// the shared-memory / L1 pipe, profiles/r01f), halves the pre-pass arithmetic of the inverse (each pair is formed
// once instead of once per element), and for P = 512 a whole frame belongs to one warp (no CTA barrier at all).
#pragma once

#include "cpq_fft.cuh"

namespace cpq
{

#ifndef CPQ_FFT16_MINB
#define CPQ_FFT16_MINB 2
#endif
template <int LOG2P>
struct Fft16Cfg
{
    static_assert(LOG2P % 3 == 0 && LOG2P >= 9, "P = 8^n, n >= 3");
    static constexpr int P = 1 << LOG2P;
    static constexpr int NB = P / 8;                 // butterflies per pass
    static constexpr int TPF = NB / 2;               // threads per frame
    static constexpr int THREADS = TPF >= 256 ? TPF : 256;
    static constexpr int FPC = THREADS / TPF;
    static constexpr int NPASS = LOG2P / 3;
    static constexpr int ROW = P + P / 8;
    static constexpr size_t SMEM = (size_t) FPC * ROW * sizeof(double2);
};

template <int TPF, int THREADS>
__device__ __forceinline__ void fft16_sync()
{
    if constexpr (TPF == 32) __syncwarp();
    else
    {
        static_assert(TPF == THREADS, "a frame is one warp or the whole CTA");
        __syncthreads();
    }
}

// spectrum-side butterfly pair of thread t and the bin of its pair r: m, and the index holding the mirror
template <int NB>
__device__ __forceinline__ int fft16_bin(int t, int r)
{
    return t != 0 ? t + r * NB : (r < 4 ? r * NB : NB / 2 + (r - 4) * NB);
}

template <int LOG2P>
__global__ void __launch_bounds__(Fft16Cfg<LOG2P>::THREADS, CPQ_FFT16_MINB) fft_fwd16_kernel(FwdArgs a)
{
    using C = Fft16Cfg<LOG2P>;
    using C8 = FftCfg<LOG2P>;
    constexpr int P = C::P, NB = C::NB, TPF = C::TPF;
    extern __shared__ double2 smem_fft[];
    const int fl = threadIdx.x / TPF;
    const int t = threadIdx.x % TPF;
    const int64_t gf = (int64_t) blockIdx.x * C::FPC + fl;
    if (gf >= a.totalFrames) return;   // a whole warp (P = 512) or the whole CTA (one frame per CTA)
    double2* buf = smem_fft + (size_t) fl * C::ROW;
    const int64_t seq = gf / a.framesPerSeq;
    const int f = (int) (gf % a.framesPerSeq);
    const double* src = a.src + seq * a.srcStride;
    const int64_t base = a.frameStart0 + (int64_t) f * P;
    const double* hist = a.histEnd ? a.histEnd + seq * a.histStride : nullptr;
    const bool inside = base >= 0 && base >= a.lo && base + 2 * P <= a.hi && ((reinterpret_cast<uintptr_t>(src + base) & 15) == 0);
    const double2* src2 = reinterpret_cast<const double2*>(src + base);
    auto gload = [&](int idx) -> double2 {
        if (inside) return __ldg(src2 + idx);
        const int64_t g = base + 2 * (int64_t) idx;
        double2 z;
        z.x = fwd_sample(src, hist, a.lo, a.hi, g);
        z.y = fwd_sample(src, hist, a.lo, a.hi, g + 1);
        return z;
    };
    auto sload = [&](int idx) -> double2 { return buf[fft_pad(idx)]; };

    double2 v[2][8];
    // ---- pass 0 (no twiddles): butterflies t and t + TPF straight from HBM ----
#pragma unroll
    for (int b = 0; b < 2; ++b) stockham_load<8, -1>(v[b], t + b * TPF, P, 1, a.ptw, gload);
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int r = 0; r < 8; ++r) buf[fft_pad(8 * (t + b * TPF) + r)] = v[b][r];
    fft16_sync<TPF, C::THREADS>();
    int Ns = 8;
#pragma unroll
    for (int p = 1; p < C::NPASS - 1; ++p)
    {
#pragma unroll
        for (int b = 0; b < 2; ++b) stockham_load<8, -1>(v[b], t + b * TPF, P, Ns, a.ptw + C8::passOffset(p), sload);
        fft16_sync<TPF, C::THREADS>();
#pragma unroll
        for (int b = 0; b < 2; ++b)
        {
            const int ob = stockham_out_base<8>(t + b * TPF, Ns);
#pragma unroll
            for (int r = 0; r < 8; ++r) buf[fft_pad(ob + r * Ns)] = v[b][r];
        }
        Ns *= 8;
        fft16_sync<TPF, C::THREADS>();
    }
    // ---- last pass (Ns = NB): butterfly t with its mirror NB - t; thread 0 takes the self-mirrored 0 and NB/2 ----
    const bool t0 = (t == 0);
    const int jB = t0 ? NB / 2 : NB - t;
    stockham_load<8, -1>(v[0], t, P, NB, a.ptw + C8::passOffset(C::NPASS - 1), sload);
    stockham_load<8, -1>(v[1], jB, P, NB, a.ptw + C8::passOffset(C::NPASS - 1), sload);
    // pair r: A[r] = Z[m_r], B[7 - r] = Z[P - m_r]   (thread 0: r = 0 is the special pair Z[0], Z[P/2])
    double2 A[8], B[8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        A[i] = v[0][i];
        A[i + 4] = t0 ? v[1][i] : v[0][i + 4];
        B[i] = t0 ? v[1][i + 4] : v[1][i];
    }
#pragma unroll
    for (int i = 4; i < 7; ++i) B[i] = t0 ? v[0][i + 1] : v[1][i];
    B[7] = t0 ? v[0][4] : v[1][7];

    // ---- split in registers: S = Z[m] + conj Z[P-m], D = Z[m] - conj Z[P-m], T = twF[m] D, X[m] = S/2 + T, X[P-m] = conj(S/2 - T) ----
    double2* out = a.out + ((size_t) seq * a.outFramesPerSeq + (size_t) (f + a.outFrameOffset)) * (size_t) P;   // packed row
    const double2* __restrict__ twF = a.tw + (P + 1);
#pragma unroll
    for (int r = 0; r < 8; ++r)
    {
        const int m = fft16_bin<NB>(t, r);
        const double2 zm = A[r], zp = B[7 - r];
        const double2 S = make_double2(zm.x + zp.x, zm.y - zp.y);
        const double2 D = make_double2(zm.x - zp.x, zm.y + zp.y);
        const double2 T = cmul(__ldg(twF + m), D);
        double2 xm = make_double2(fma(0.5, S.x, T.x), fma(0.5, S.y, T.y));
        double2 xp = make_double2(fma(0.5, S.x, -T.x), fma(-0.5, S.y, T.y));
        int mp = P - m;
        if (r == 0 && t0)
        {
            xm = make_double2(zm.x + zm.y, zm.x - zm.y);   // packed slot 0 = (Re X[0], Re X[P])
            xp = make_double2(zp.x, -zp.y);                // X[P/2] = conj Z[P/2]
            mp = P / 2;
        }
        out[m] = xm;
        out[mp] = xp;
    }
}

template <int LOG2P>
__global__ void __launch_bounds__(Fft16Cfg<LOG2P>::THREADS, CPQ_FFT16_MINB) fft_inv16_kernel(InvArgs a)
{
    using C = Fft16Cfg<LOG2P>;
    using C8 = FftCfg<LOG2P>;
    constexpr int P = C::P, NB = C::NB, TPF = C::TPF;
    extern __shared__ double2 smem_fft[];
    const int fl = threadIdx.x / TPF;
    const int t = threadIdx.x % TPF;
    const int64_t gf = (int64_t) blockIdx.x * C::FPC + fl;
    if (gf >= a.totalFrames) return;
    double2* buf = smem_fft + (size_t) fl * C::ROW;
    const int64_t seq = gf / a.framesOut;
    const int f = (int) (gf % a.framesOut);
    const double2* __restrict__ Y = a.in + ((size_t) seq * a.framesPerSeq + (size_t) f) * (size_t) P;   // packed row
    const double2* __restrict__ twI = a.tw + 2 * (P + 1);
    const double invN = 1.0 / (double) (2 * P);
    auto sload = [&](int idx) -> double2 { return buf[fft_pad(idx)]; };

    // ---- pre-pass in registers: Z[m] = S/N + T, Z[P-m] = conj(S/N - T), S = Y[m] + conj Y[P-m], D = Y[m] - conj Y[P-m],
    // T = twI[m] D; pair r of this thread feeds element r of butterfly t and element 7 - r of butterfly NB - t ----
    const bool t0 = (t == 0);
    double2 A[8], B[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
    {
        const int m = fft16_bin<NB>(t, r);
        const bool special = (r == 0) && t0;
        const double2 ym = __ldg(Y + m);
        const double2 yp = __ldg(Y + (special ? P / 2 : P - m));
        const double2 S = make_double2(ym.x + yp.x, ym.y - yp.y);
        const double2 D = make_double2(ym.x - yp.x, ym.y + yp.y);
        const double2 T = cmul(__ldg(twI + m), D);
        double2 za = make_double2(fma(S.x, invN, T.x), fma(S.y, invN, T.y));
        double2 zb = make_double2(fma(S.x, invN, -T.x), fma(-S.y, invN, T.y));
        if (special)
        {
            za = make_double2((ym.x + ym.y) * invN, (ym.x - ym.y) * invN);   // packed (Re Y[0], Re Y[P])
            zb = make_double2(2.0 * invN * yp.x, -2.0 * invN * yp.y);        // Z[P/2] = 2 conj Y[P/2] / N
        }
        A[r] = za;
        B[7 - r] = zb;
    }
    double2 v[2][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        v[0][i] = A[i];
        v[1][i] = t0 ? A[i + 4] : B[i];
        v[1][i + 4] = t0 ? B[i] : B[i + 4];
    }
    v[0][4] = t0 ? B[7] : A[4];
#pragma unroll
    for (int i = 5; i < 8; ++i) v[0][i] = t0 ? B[i - 1] : A[i];
    // ---- pass 0 (no twiddles) ----
    const int jB = t0 ? NB / 2 : NB - t;
    dft8<+1>(v[0]);
    dft8<+1>(v[1]);
#pragma unroll
    for (int r = 0; r < 8; ++r)
    {
        buf[fft_pad(8 * t + r)] = v[0][r];
        buf[fft_pad(8 * jB + r)] = v[1][r];
    }
    fft16_sync<TPF, C::THREADS>();
    int Ns = 8;
#pragma unroll
    for (int p = 1; p < C::NPASS - 1; ++p)
    {
#pragma unroll
        for (int b = 0; b < 2; ++b) stockham_load<8, +1>(v[b], t + b * TPF, P, Ns, a.ptw + C8::passOffset(p), sload);
        fft16_sync<TPF, C::THREADS>();
#pragma unroll
        for (int b = 0; b < 2; ++b)
        {
            const int ob = stockham_out_base<8>(t + b * TPF, Ns);
#pragma unroll
            for (int r = 0; r < 8; ++r) buf[fft_pad(ob + r * Ns)] = v[b][r];
        }
        Ns *= 8;
        fft16_sync<TPF, C::THREADS>();
    }
    // ---- last pass: outputs z[j + r NB]; only z[P/2 ..) = y[P .. 2P) is kept and goes straight to HBM ----
    double2* o = reinterpret_cast<double2*>(a.out + seq * a.outStride + (int64_t) f * P);
#pragma unroll
    for (int b = 0; b < 2; ++b)
    {
        const int j = t + b * TPF;
        stockham_load<8, +1>(v[b], j, P, NB, a.ptw + C8::passOffset(C::NPASS - 1), sload);
#pragma unroll
        for (int r = 4; r < 8; ++r) o[j + (r - 4) * NB] = v[b][r];
    }
}

} // namespace cpq
