// cpq_mac.cuh -- frequency-domain delay-line multiply-accumulate (sm_100a, FP64).
//
//   Y[k][m] = sum_{q in [qBegin,qEnd)} X[k-q][m] * H[q][m]
//
// is the MAC of processLayerBlock / Add (MKLNonUniformConvolver.cpp:1293-1308, 1505-1520;
// accumulateSplitComplex :150-195) for every frame k at once: a Q-tap complex FIR along the frame index,
// independently per bin.  Frames with k-q < 0 are the zero history of a Reset engine.  The reference's FDL
// ring / mirror slots / reversed partition order / partsPerCallback time slicing only change *when* the
// products are formed, not their sum.
//
// Spectra are stored "packed": P complex slots per frame, slot 0 = (Re X[0], Re X[P]) -- bins 0 and P of a real
// signal's spectrum are real -- so a row is exactly P * 16 bytes and every 64-bin tile is full.
//
// One CTA owns 64 bins of one sequence and a range of frames.  The IR spectra tile H[q][64 bins] is staged in
// shared memory once (cp.async) and reused for every frame; four thread groups work on four runs of KT = 8
// consecutive output frames at a time.  Each thread keeps its 8 accumulators and a sliding window of 8 input
// spectra in registers: per tap it loads one H value (shared memory) and one new X value (global, one step
// ahead), so the FP64 pipe sees 32 DFMAs per two 16-byte loads.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace cpq
{

struct MacArgs
{
    const double2* X;    // [nSeq][K][P]
    const double2* H;    // [nH][Q][P]   natural partition order q = 0..Q-1
    double2* Y;          // [nSeq][K][P]
    int K, P, Q;
    int qBegin, qEnd;
    int64_t hSeqStride;  // elements between sequences in H
    int hSeqMod;         // H row = seq % hSeqMod when the IR pair is shared by all streams; 0 = seq
    int framesPerCta;    // multiple of 4*KT
};

constexpr int kMacBins = 64;
constexpr int kMacGroups = 4;
constexpr int kMacThreads = kMacBins * kMacGroups;
constexpr int kMacKT = 8;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(kMacThreads, 2) mac_kernel(MacArgs a)
{
    constexpr int KT = kMacKT;
    extern __shared__ __align__(16) double2 Hs[];   // [qEnd - qBegin][64]
    const int ml = threadIdx.x & (kMacBins - 1);
    const int g = threadIdx.x / kMacBins;
    const int m = blockIdx.x * kMacBins + ml;        // P is a multiple of 64: always in range
    const int seq = blockIdx.z;
    const int kc0 = blockIdx.y * a.framesPerCta;
    const int kc1 = min(a.K, kc0 + a.framesPerCta);
    const int nq = a.qEnd - a.qBegin;
    const int hrow = a.hSeqMod > 0 ? (seq % a.hSeqMod) : seq;

    // ---- stage the IR spectra tile ----
    {
        const double2* __restrict__ H = a.H + (size_t) hrow * a.hSeqStride + (size_t) a.qBegin * a.P + (size_t) blockIdx.x * kMacBins;
        for (int i = threadIdx.x; i < nq * kMacBins; i += kMacThreads)
            cp_async16(Hs + i, H + (size_t) (i / kMacBins) * a.P + (i & (kMacBins - 1)));
        cp_async_wait_all();
    }
    __syncthreads();

    const double2* __restrict__ X = a.X + (size_t) seq * a.K * a.P + m;
    double2* __restrict__ Y = a.Y + (size_t) seq * a.K * a.P + m;
    const bool packed = (m == 0);   // slot 0 holds two real bins: (re*re, im*im) instead of a complex product

    auto loadX = [&](int f) -> double2 { return (f >= 0 && f < a.K) ? __ldg(X + (size_t) f * a.P) : make_double2(0.0, 0.0); };

    for (int ks = kc0 + g * KT; ks < kc1; ks += kMacGroups * KT)
    {
        double2 acc[KT], w[KT];
#pragma unroll
        for (int i = 0; i < KT; ++i)
        {
            acc[i] = make_double2(0.0, 0.0);
            w[i] = loadX(ks + i - a.qBegin);   // logical window for tap qBegin
        }
        double2 nxt = loadX(ks - a.qBegin - 1);   // next older frame, one tap ahead
        for (int q0 = 0; q0 < nq; q0 += KT)
        {
#pragma unroll
            for (int u = 0; u < KT; ++u)
            {
                const int q = q0 + u;
                if (q < nq)   // uniform
                {
                    const double2 h = Hs[q * kMacBins + ml];
                    const double hA = h.x, hB = packed ? 0.0 : -h.y, hC = packed ? 0.0 : h.y, hD = packed ? h.y : h.x;
                    const double2 incoming = nxt;
                    nxt = loadX(ks - (a.qBegin + q) - 2);
#pragma unroll
                    for (int i = 0; i < KT; ++i)
                    {
                        const double2 x = w[(i + KT - u) % KT];   // logical w[i] at tap q
                        acc[i].x = fma(x.x, hA, fma(x.y, hB, acc[i].x));
                        acc[i].y = fma(x.x, hC, fma(x.y, hD, acc[i].y));
                    }
                    // slide: logical w[i] <- w[i-1], w[0] <- frame ks - q - 1; the freed physical slot is (KT-1-u)
                    w[(KT - 1 - u) % KT] = incoming;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < KT; ++i)
            if (ks + i < kc1) Y[(size_t) (ks + i) * a.P] = acc[i];
    }
}

} // namespace cpq
