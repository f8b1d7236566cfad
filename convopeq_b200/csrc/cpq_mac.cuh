// cpq_mac.cuh -- frequency-domain delay-line multiply-accumulate (sm_100a, FP64).
//
//   Y[k][m] = sum_{q in [qBegin,qEnd)} X[k-q][m] * H[q][m]
//
// is the MAC of processLayerBlock / Add (MKLNonUniformConvolver.cpp:1293-1308, 1505-1520;
// accumulateSplitComplex :150-195) for every frame k at once: a Q-tap complex FIR along the frame index,
// independently per bin.  Frames with k-q < 0 are the zero history of a Reset engine.  The reference's FDL
// ring / mirror slots / reversed partition order / partsPerCallback time slicing only change *when* the
// products are formed, not their sum.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace cpq
{

struct MacArgs
{
    const double2* X;    // [nSeq][K][M]
    const double2* H;    // [nH][Q][M]   natural partition order q = 0..Q-1
    double2* Y;          // [nSeq][K][M]
    int K, M, Q;
    int qBegin, qEnd;
    int64_t hSeqStride;  // elements between sequences in H
    int hSeqMod;         // H row = seq % hSeqMod when the IR pair is shared by all streams; 0 = seq
};

// Register-blocked KT outputs x QT taps per thread, one thread per bin.
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

} // namespace cpq
