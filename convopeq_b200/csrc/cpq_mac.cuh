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
// signal's spectrum are real -- so a row is exactly P * 16 bytes and every 32-bin tile is full.
//
// One CTA owns 32 bins of one sequence and a range of frames.  Both operands are staged in shared memory with
// cp.async: the IR spectra tile H[q][32] once, and the input spectra X[f][32] through a ring of rows that is
// refilled one super-step (64 output frames) ahead of the arithmetic, so each X and H element is fetched from
// HBM/L2 once per CTA and no FP64 warp ever waits on a global load.  Eight thread groups work on eight runs of
// KT = 8 consecutive output frames; each thread keeps 8 accumulators and a sliding window of 8 input spectra
// in registers and per tap reads one H and one new X value from shared memory: 32 DFMAs per two LDS.128.
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
    int hSeqMod;         // H row = (seqBase + seq) % hSeqMod when the IR pair is shared by all streams; 0 = seq
    int seqBase;         // absolute index of this launch's first sequence (sequence chunks)
    int framesPerCta;    // multiple of kMacSuper
    int ringRows;        // >= 2*kMacSuper + (qEnd - qBegin) - 1
};

constexpr int kMacBins = 32;
constexpr int kMacGroups = 8;
constexpr int kMacThreads = kMacBins * kMacGroups;   // 256
constexpr int kMacKT = 8;
constexpr int kMacSuper = kMacGroups * kMacKT;       // 64 output frames per super-step

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

inline size_t macSmemBytes(int nq, int ringRows) { return (size_t) (nq + ringRows) * kMacBins * sizeof(double2); }
inline int macRingRows(int nq) { return 2 * kMacSuper + nq - 1; }

// PACKED = this CTA's tile contains slot 0, which holds two real bins: its product is (re*re, im*im) instead of a
// complex product.  Only the first bin tile pays for the operand selects.
template <bool PACKED>
__device__ __forceinline__ void mac_body(const MacArgs& a, double2* __restrict__ Hs, double2* __restrict__ ring, int nq, int R,
                                         int ml, int g, int kc0, int kc1, double2* __restrict__ Y, const double2* __restrict__ X,
                                         int m0)
{
    constexpr int KT = kMacKT;
    const int rowBytes = kMacBins * (int) sizeof(double2);
    const int ringBytes = R * rowBytes;
    const char* ringB = reinterpret_cast<const char*>(ring) + ml * (int) sizeof(double2);
    const char* hsB = reinterpret_cast<const char*>(Hs) + ml * (int) sizeof(double2);
    const bool slot0 = PACKED && (m0 + ml == 0);

    auto slotOf = [&](int f) -> int { int s = f % R; return s < 0 ? s + R : s; };
    auto fillRows = [&](int f0, int f1) {
        const int n = (f1 - f0) * kMacBins;
        for (int i = threadIdx.x; i < n; i += kMacThreads)
        {
            const int f = f0 + i / kMacBins, c = i & (kMacBins - 1);
            double2* dst = ring + slotOf(f) * kMacBins + c;
            if (f >= 0 && f < a.K) cp_async16(dst, X + (size_t) f * a.P + c);
            else *dst = make_double2(0.0, 0.0);
        }
    };

    for (int ks0 = kc0; ks0 < kc1; ks0 += kMacSuper)
    {
        // prefetch the rows that only the next super-step needs
        if (ks0 + kMacSuper < kc1)
        {
            fillRows(ks0 - a.qBegin + kMacSuper, ks0 - a.qBegin + 2 * kMacSuper);
            cp_async_commit();
        }
        const int ks = ks0 + g * KT;
        if (ks < kc1)
        {
            double2 acc[KT], w[KT];
            int off = slotOf(ks - a.qBegin) * rowBytes;   // byte offset of frame ks - qBegin in the ring
#pragma unroll
            for (int i = 0; i < KT; ++i)
            {
                acc[i] = make_double2(0.0, 0.0);
                int o = off + i * rowBytes;
                if (o >= ringBytes) o -= ringBytes;
                w[i] = *reinterpret_cast<const double2*>(ringB + o);   // logical window for the first tap: frame ks + i - qBegin
            }
            off -= rowBytes;
            if (off < 0) off += ringBytes;
            double2 nxt = *reinterpret_cast<const double2*>(ringB + off);   // frame ks - qBegin - 1
            for (int q0 = 0; q0 < nq; q0 += KT)
            {
#pragma unroll
                for (int u = 0; u < KT; ++u)
                {
                    const int q = q0 + u;
                    if (q < nq)   // uniform
                    {
                        const double2 h = *reinterpret_cast<const double2*>(hsB + q * rowBytes);
                        const double2 incoming = nxt;
                        off -= rowBytes;
                        if (off < 0) off += ringBytes;
                        nxt = *reinterpret_cast<const double2*>(ringB + off);   // frame ks - qBegin - q - 2 (unused after the last tap)
                        if (PACKED)
                        {
                            const double hA = h.x, hB = slot0 ? 0.0 : -h.y, hC = slot0 ? 0.0 : h.y, hD = slot0 ? h.y : h.x;
#pragma unroll
                            for (int i = 0; i < KT; ++i)
                            {
                                const double2 x = w[(i + KT - u) % KT];
                                acc[i].x = fma(x.x, hA, fma(x.y, hB, acc[i].x));
                                acc[i].y = fma(x.x, hC, fma(x.y, hD, acc[i].y));
                            }
                        }
                        else
                        {
                            const double nhy = -h.y;
#pragma unroll
                            for (int i = 0; i < KT; ++i)
                            {
                                const double2 x = w[(i + KT - u) % KT];   // logical w[i] at tap q
                                acc[i].x = fma(x.x, h.x, fma(x.y, nhy, acc[i].x));
                                acc[i].y = fma(x.x, h.y, fma(x.y, h.x, acc[i].y));
                            }
                        }
                        // slide: logical w[i] <- w[i-1], w[0] <- frame ks - qBegin - q - 1; freed physical slot is (KT-1-u)
                        w[(KT - 1 - u) % KT] = incoming;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < KT; ++i)
                if (ks + i < kc1) Y[(size_t) (ks + i) * a.P] = acc[i];
        }
        cp_async_wait0();
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kMacThreads, 2) mac_kernel(MacArgs a)
{
    extern __shared__ __align__(16) double2 mac_smem[];
    const int nq = a.qEnd - a.qBegin;
    const int R = a.ringRows;
    double2* Hs = mac_smem;                       // [nq][32]
    double2* ring = mac_smem + nq * kMacBins;     // [R][32], row of frame f lives in slot f mod R
    const int ml = threadIdx.x & (kMacBins - 1);
    const int g = threadIdx.x / kMacBins;
    const int m0 = blockIdx.x * kMacBins;          // P is a multiple of 32: tiles are always full
    const int seq = blockIdx.z;
    const int kc0 = blockIdx.y * a.framesPerCta;
    const int kc1 = min(a.K, kc0 + a.framesPerCta);
    const int hrow = a.hSeqMod > 0 ? ((a.seqBase + seq) % a.hSeqMod) : seq;
    const double2* __restrict__ X = a.X + (size_t) seq * a.K * a.P + m0;
    double2* __restrict__ Y = a.Y + (size_t) seq * a.K * a.P + m0 + ml;

    // ---- stage the IR spectra tile and the first super-step's input rows ----
    {
        const double2* __restrict__ H = a.H + (size_t) hrow * a.hSeqStride + (size_t) a.qBegin * a.P + m0;
        for (int i = threadIdx.x; i < nq * kMacBins; i += kMacThreads)
            cp_async16(Hs + i, H + (size_t) (i / kMacBins) * a.P + (i & (kMacBins - 1)));
        const int f0 = kc0 - a.qBegin - (nq - 1), f1 = kc0 - a.qBegin + kMacSuper;
        const int n = (f1 - f0) * kMacBins;
        for (int i = threadIdx.x; i < n; i += kMacThreads)
        {
            const int f = f0 + i / kMacBins, c = i & (kMacBins - 1);
            int s = f % R;
            if (s < 0) s += R;
            double2* dst = ring + s * kMacBins + c;
            if (f >= 0 && f < a.K) cp_async16(dst, X + (size_t) f * a.P + c);
            else *dst = make_double2(0.0, 0.0);
        }
    }
    cp_async_commit();
    cp_async_wait0();
    __syncthreads();

    if (m0 == 0) mac_body<true>(a, Hs, ring, nq, R, ml, g, kc0, kc1, Y, X, m0);
    else mac_body<false>(a, Hs, ring, nq, R, ml, g, kc0, kc1, Y, X, m0);
}

} // namespace cpq
