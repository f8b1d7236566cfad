// cpq_fft_large.cuh -- real FFTs whose complex length P exceeds one CTA's shared memory (P = 16384 .. 65536,
// i.e. the reference's 32768 / 65536 / 131072-point layer-2 transforms, FFTBackend.cpp:123-151).
//
// Same mathematics as cpq_fft.cuh (P-point complex Stockham autosort FFT of z[n] = x[2n] + i x[2n+1] plus a
// split pass, packed CCS rows), but every radix pass is its own kernel that ping-pongs between two global
// buffers.  Only the tail layer of very long IRs (cfg 5: 2M taps, 56 partitions of 32768) takes this path; it
// moves each frame through HBM once per pass, which is acceptable for the handful of frames involved.
#pragma once

#include "cpq_fft.cuh"

namespace cpq
{

// W_P^idx from the layer table tw[t] = exp(-2 pi i t / (2P)), t = 0..P; SIGN > 0 conjugates.
template <int SIGN>
__device__ __forceinline__ double2 twiddleFromLayerTable(const double2* __restrict__ tw, int P, int64_t idx)
{
    const int64_t t = 2 * idx;
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

struct LargeFftArgs
{
    int P;                  // complex length
    int64_t totalFrames;
    int framesPerSeq;
    // real side
    const double* src;      // forward: [nSeq][srcStride]
    int64_t srcStride, frameStart0, lo, hi;
    int halfOnly;
    double* dst;            // inverse: [nSeq][dstStride], frame f -> dst[f*P .. (f+1)*P)
    int64_t dstStride;
    // complex buffers, rows of P double2 per frame
    double2* a;
    double2* b;
    int rowPitchFrames;     // frames per sequence in a/b
    int rowOffset;          // frame f of a sequence lives in row f + rowOffset
    const double2* tw;
    double scale;
    int applyScale;
    const double* gain;
    const double* tilt;
};

__device__ __forceinline__ size_t largeRow(const LargeFftArgs& a, int64_t gf)
{
    const int64_t seq = gf / a.framesPerSeq, f = gf % a.framesPerSeq;
    return ((size_t) seq * a.rowPitchFrames + (size_t) (f + a.rowOffset)) * (size_t) a.P;
}

// z[idx] = x[2 idx] + i x[2 idx + 1] of frame gf -> a
__global__ void gfft_load_fwd_kernel(LargeFftArgs a)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.totalFrames * a.P) return;
    const int64_t gf = i / a.P;
    const int idx = (int) (i % a.P);
    const int64_t seq = gf / a.framesPerSeq, f = gf % a.framesPerSeq;
    const double* src = a.src + seq * a.srcStride;
    const int64_t g = a.frameStart0 + f * (int64_t) a.P + 2 * (int64_t) idx;
    double2 z = make_double2(0.0, 0.0);
    if (!(a.halfOnly && 2 * idx >= a.P))
    {
        z.x = (g >= a.lo && g < a.hi) ? __ldg(src + g) : 0.0;
        z.y = (g + 1 >= a.lo && g + 1 < a.hi) ? __ldg(src + g + 1) : 0.0;
    }
    a.a[largeRow(a, gf) + idx] = z;
}

// one Stockham pass, in -> out (both rows of P), sub-transform length Ns before the pass
template <int R, int SIGN>
__global__ void gfft_pass_kernel(LargeFftArgs a, const double2* __restrict__ in, double2* __restrict__ out, int Ns)
{
    const int per = a.P / R;
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.totalFrames * per) return;
    const int64_t gf = i / per;
    const int j = (int) (i % per);
    const size_t row = largeRow(a, gf);
    const int k = j & (Ns - 1);
    const int64_t tscale = a.P / ((int64_t) Ns * R);
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
    {
        double2 x = in[row + j + (size_t) r * per];
        if (r > 0 && Ns > 1) x = cmul(x, twiddleFromLayerTable<SIGN>(a.tw, a.P, (int64_t) k * r * tscale));
        v[r] = x;
    }
    dftR<R, SIGN>(v);
    const int ob = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) out[row + ob + (size_t) r * Ns] = v[r];
}

// in-place split pass on buffer z -> packed CCS (thread owns the pair m, P-m)
__global__ void gfft_split_fwd_kernel(LargeFftArgs a, double2* __restrict__ z)
{
    const int half = a.P / 2 + 1;
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.totalFrames * half) return;
    const int64_t gf = i / half;
    const int m = (int) (i % half);
    const int P = a.P;
    double2* row = z + largeRow(a, gf);
    const double2 zm = row[m];
    const double2 zc = cconj(row[(P - m) & (P - 1)]);
    const double2 E = make_double2(0.5 * (zm.x + zc.x), 0.5 * (zm.y + zc.y));
    const double2 D = make_double2(0.5 * (zm.x - zc.x), 0.5 * (zm.y - zc.y));
    const double2 O = make_double2(D.y, -D.x);
    const double2 Tm = cmul(__ldg(a.tw + m), O);
    auto fin = [&](int bin, double2 X) -> double2 {
        if (a.applyScale) { X.x *= a.scale; X.y *= a.scale; }
        if (a.gain) { const double g = __ldg(a.gain + bin); X.x *= g; X.y *= g; }
        if (a.tilt) { const double g = __ldg(a.tilt + bin); X.x *= g; X.y *= g; }
        return X;
    };
    if (m == 0)
    {
        const double2 x0 = fin(0, make_double2(E.x + O.x, 0.0));
        const double2 xP = fin(P, make_double2(E.x - O.x, 0.0));
        row[0] = make_double2(x0.x, xP.x);
    }
    else
    {
        const double2 xm = fin(m, cadd(E, Tm));
        if (m != P / 2)
        {
            const double2 xc = fin(P - m, cconj(csub(E, Tm)));
            row[P - m] = xc;
        }
        row[m] = xm;
    }
}

// in-place inverse pre-pass: packed CCS Y -> Z (thread owns the pair m, P-m)
__global__ void gfft_pre_inv_kernel(LargeFftArgs a, double2* __restrict__ y)
{
    const int half = a.P / 2 + 1;
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.totalFrames * half) return;
    const int64_t gf = i / half;
    const int m = (int) (i % half);
    const int P = a.P;
    const double invN = 1.0 / (double) (2 * P);
    double2* row = y + largeRow(a, gf);
    auto zOf = [&](double2 ym, double2 yc, int bin) -> double2 {
        const double2 S = cadd(ym, yc), D = csub(ym, yc);
        const double2 Tm = cmul(cconj(__ldg(a.tw + bin)), D);
        return make_double2((S.x - Tm.y) * invN, (S.y + Tm.x) * invN);
    };
    if (m == 0)
    {
        const double2 y0 = row[0];
        row[0] = zOf(make_double2(y0.x, 0.0), make_double2(y0.y, 0.0), 0);
    }
    else
    {
        const double2 ym = row[m], yp = row[P - m];
        const double2 zm = zOf(ym, cconj(yp), m);
        if (m != P / 2) row[P - m] = zOf(yp, cconj(ym), P - m);
        row[m] = zm;
    }
}

// z[P/2 ..) -> real samples y[P .. 2P) of each frame
__global__ void gfft_store_inv_kernel(LargeFftArgs a, const double2* __restrict__ z)
{
    const int half = a.P / 2;
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.totalFrames * half) return;
    const int64_t gf = i / half;
    const int n = (int) (i % half);
    const int64_t seq = gf / a.framesPerSeq, f = gf % a.framesPerSeq;
    double2* o = reinterpret_cast<double2*>(a.dst + seq * a.dstStride + f * (int64_t) a.P);
    o[n] = z[largeRow(a, gf) + half + n];
}

} // namespace cpq
