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
    const double* histEnd;  // see FwdArgs
    int64_t histStride;
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
    const double* hist = a.histEnd ? a.histEnd + seq * a.histStride : nullptr;
    const int64_t g = a.frameStart0 + f * (int64_t) a.P + 2 * (int64_t) idx;
    double2 z = make_double2(0.0, 0.0);
    if (!(a.halfOnly && 2 * idx >= a.P))
    {
        z.x = fwd_sample(src, hist, a.lo, a.hi, g);
        z.y = fwd_sample(src, hist, a.lo, a.hi, g + 1);
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


// ---------------------------------------------------------------------------------------------
// Four-step form of the same P-point complex transform (round 2): P = N1 x 256, index n = n1 * 256 + n2 in, k = k1 + N1 * k2
// out.  Two kernels, each moving a frame through HBM once:
//   cols  for 16 consecutive n2 the N1-point transforms along n1 (stride 256) in shared memory, then the twiddle
//         W_P^(k1 n2); element (k1, n2) goes back to position k1 * 256 + n2 (same places: a column block is private to its CTA)
//   rows  for 8 consecutive k1 the 256-point transforms along n2 (contiguous rows, one warp each); element (k1, k2) goes to
//         k1 + N1 * k2, eight consecutive k1 per 128-byte run
// instead of one kernel per radix-8 pass (five of them at P = 32768, each reading and writing the whole frame, plus a load
// pass): the forward transform reads the signal directly in `cols` (the complex view of the real frame is a reinterpretation),
// the inverse writes the kept half of each frame straight to the real output in `rows`.  Twiddles of every size come from the
// layer table (W_P^idx), so the small transforms add no tables.
// ---------------------------------------------------------------------------------------------
constexpr int kG2Cols = 32;       // columns per CTA in the column pass (512-byte runs per row of the frame)
constexpr int kG2Rows = 8;        // rows per CTA in the row pass
constexpr int kG2N2 = 256;

// one Stockham pass of an N-point transform held in shared memory (stride 1), NT threads per transform, thread t.
// All threads of the CTA call it together: `sync` separates the reads of a pass from its writes.
// wN[m] = W_N^m (already conjugated for the inverse), a table in shared memory built once per CTA from the layer table
template <int N, int NT, int R, int SIGN, class SyncF>
__device__ __forceinline__ void g2_pass(double2* s, int t, int Ns, const double2* __restrict__ wN, SyncF sync)
{
    constexpr int per = N / R;
    constexpr int cnt = per / NT;          // butterflies per thread
    static_assert(per % NT == 0 && cnt >= 1, "pass geometry");
    double2 v[cnt][R];
#pragma unroll
    for (int c = 0; c < cnt; ++c)
    {
        const int j = t + c * NT;
        const int k = j & (Ns - 1);
#pragma unroll
        for (int r = 0; r < R; ++r)
        {
            double2 x = s[j + r * per];
            if (r > 0 && Ns > 1) x = cmul(x, wN[k * r * (N / (Ns * R))]);   // W_{Ns R}^{k r},  k r < Ns R
            v[c][r] = x;
        }
        dftR<R, SIGN>(v[c]);
    }
    sync();
#pragma unroll
    for (int c = 0; c < cnt; ++c)
    {
        const int j = t + c * NT;
        const int k = j & (Ns - 1);
        const int ob = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) s[ob + r * Ns] = v[c][r];
    }
    sync();
}

// N-point transform in shared memory, N in {64, 128, 256}, N / 8 threads
template <int N, int SIGN, class SyncF>
__device__ __forceinline__ void g2_fft(double2* s, int t, const double2* __restrict__ wN, SyncF sync)
{
    constexpr int NT = N / 8;
    int Ns = 1;
    if constexpr (N == 128) { g2_pass<N, NT, 2, SIGN>(s, t, Ns, wN, sync); Ns = 2; }
    if constexpr (N == 256) { g2_pass<N, NT, 4, SIGN>(s, t, Ns, wN, sync); Ns = 4; }
    g2_pass<N, NT, 8, SIGN>(s, t, Ns, wN, sync);
    Ns *= 8;
    g2_pass<N, NT, 8, SIGN>(s, t, Ns, wN, sync);
}

template <int LOG2N1, int SIGN, bool FROM_REAL>
__global__ void __launch_bounds__(kG2Cols * (1 << LOG2N1) / 8) gfft2_cols_kernel(LargeFftArgs a, const double2* __restrict__ in, double2* __restrict__ out)
{
    constexpr int N1 = 1 << LOG2N1, C = kG2Cols, LD = N1 + 1, NT = N1 / 8, THREADS = C * NT;
    extern __shared__ double2 g2_smem[];
    const int tid = threadIdx.x;
    const int64_t gf = blockIdx.x / (kG2N2 / C);
    const int n20 = (int) (blockIdx.x % (kG2N2 / C)) * C;
    const size_t row = largeRow(a, gf);
    const int P = a.P;
    const int64_t seq = gf / a.framesPerSeq, f = gf % a.framesPerSeq;
    const double* src = FROM_REAL ? a.src + seq * a.srcStride : nullptr;
    const double* hist = (FROM_REAL && a.histEnd) ? a.histEnd + seq * a.histStride : nullptr;
    const int64_t g0 = FROM_REAL ? a.frameStart0 + f * (int64_t) P : 0;
    for (int idx = tid; idx < N1 * C; idx += THREADS)
    {
        const int n1 = idx / C, c = idx % C;
        const int e = n1 * kG2N2 + n20 + c;
        double2 z;
        if (FROM_REAL)
        {
            z = make_double2(0.0, 0.0);
            if (!(a.halfOnly && 2 * e >= P))
            {
                const int64_t g = g0 + 2 * (int64_t) e;
                z.x = fwd_sample(src, hist, a.lo, a.hi, g);
                z.y = fwd_sample(src, hist, a.lo, a.hi, g + 1);
            }
        }
        else
            z = in[row + e];
        g2_smem[c * LD + n1] = z;
    }
    double2* wN = g2_smem + C * LD;                        // W_N1^m, m < N1
    for (int m = tid; m < N1; m += THREADS) wN[m] = twiddleFromLayerTable<SIGN>(a.tw, P, (int64_t) m * (P / N1));
    __syncthreads();
    g2_fft<N1, SIGN>(g2_smem + (tid / NT) * LD, tid % NT, wN, [] { __syncthreads(); });
    for (int idx = tid; idx < N1 * C; idx += THREADS)
    {
        const int k1 = idx / C, c = idx % C;
        const double2 v = cmul(g2_smem[c * LD + k1], twiddleFromLayerTable<SIGN>(a.tw, P, (int64_t) k1 * (n20 + c)));
        out[row + (size_t) k1 * kG2N2 + n20 + c] = v;
    }
}

template <int LOG2N1, int SIGN, bool TO_REAL>
__global__ void __launch_bounds__(kG2Rows * 32) gfft2_rows_kernel(LargeFftArgs a, const double2* __restrict__ in, double2* __restrict__ out)
{
    constexpr int N1 = 1 << LOG2N1, R = kG2Rows, LD = kG2N2 + 1;
    extern __shared__ double2 g2_smem[];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int64_t gf = blockIdx.x / (N1 / R);
    const int k10 = (int) (blockIdx.x % (N1 / R)) * R;
    const size_t row = largeRow(a, gf);
    const int P = a.P;
    double2* mine = g2_smem + w * LD;
    const double2* rin = in + row + (size_t) (k10 + w) * kG2N2;
#pragma unroll
    for (int i = 0; i < kG2N2 / 32; ++i) mine[lane + 32 * i] = rin[lane + 32 * i];
    double2* wN = g2_smem + R * LD;                        // W_256^m
    wN[tid] = twiddleFromLayerTable<SIGN>(a.tw, P, (int64_t) tid * (P / kG2N2));   // R * 32 = 256 threads
    __syncthreads();
    g2_fft<kG2N2, SIGN>(mine, lane, wN, [] { __syncwarp(); });
    __syncthreads();
    const int64_t seq = gf / a.framesPerSeq, f = gf % a.framesPerSeq;
    for (int idx = tid; idx < kG2N2 * R; idx += R * 32)
    {
        const int k2 = idx / R, rr = idx % R;
        const double2 v = g2_smem[rr * LD + k2];
        const int n = (k10 + rr) + N1 * k2;
        if (TO_REAL)
        {
            // z[P/2 ..) are the samples [P, 2P) of the frame: the half overlap-save keeps
            if (n >= P / 2) reinterpret_cast<double2*>(a.dst + seq * a.dstStride + f * (int64_t) P)[n - P / 2] = v;
        }
        else
            out[row + n] = v;
    }
}

} // namespace cpq
