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
// One CTA owns 32 bins of one sequence and a range of frames.  Both operands are staged in shared memory by the
// TMA engine (cp.async.bulk, one 512-byte row per copy, completion on an mbarrier): the IR spectra tile H[q][32]
// once, and the input spectra X[f][32] through a ring of rows that is refilled one super-step (64 output frames)
// ahead of the arithmetic (eight rows per warp), so each X and H element is fetched from HBM/L2 once per CTA, no FP64 warp
// ever waits on a global load and no thread spends instructions on per-element copy addressing.  Eight thread
// groups work on eight runs of KT = 8 consecutive output frames; each thread keeps 8 accumulators and a sliding
// window of 8 input spectra in registers and per tap reads one H and one new X value from shared memory: 32 DFMAs
// per two LDS.128.  Taps are consumed in fully unrolled blocks of 8 (one full rotation of the register window, so
// the rotation costs no moves) plus one compile-time-sized remainder block.
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
    int hist;            // rows of X stored before frame 0 of each sequence (streaming continuation: the carried FDL, frames
                         // -hist..-1); frames below -hist are the zero history of a Reset engine.  0 in the one-shot form
    int xRows;           // rows per sequence in X = hist + K
};

constexpr int kMacBins = 32;
constexpr int kMacGroups = 8;
constexpr int kMacThreads = kMacBins * kMacGroups;   // 256
constexpr int kMacKT = 8;
constexpr int kMacSuper = kMacGroups * kMacKT;       // 64 output frames per super-step
constexpr int kMacRowBytes = kMacBins * (int) sizeof(double2);   // 512

inline size_t macSmemBytes(int nq, int ringRows) { return (size_t) (nq + ringRows) * kMacRowBytes + 16; }
inline int macRingRows(int nq) { return (2 * kMacSuper + nq - 1 + kMacKT - 1) / kMacKT * kMacKT; }   // a multiple of KT: mac_taps3<.., ALIGNED>

// ---- mbarrier / bulk-copy (TMA engine) primitives ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dstSmem, const void* srcGlobal, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dstSmem)),
                 "l"(srcGlobal), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// NT consecutive taps starting at H row q0, window in canonical order at entry (w[i] = frame ks + i - qBegin - q0,
// nxt = the frame before w[0]); canonical again on exit iff NT == KT.
// PACKED = this CTA's tile contains slot 0, which holds two real bins: its product is (re*re, im*im) instead of a
// complex product.  Only the first bin tile pays for the operand selects.
template <int NT, bool PACKED>
__device__ __forceinline__ void mac_taps(double2 (&acc)[kMacKT], double2 (&w)[kMacKT], double2& nxt, const char* __restrict__ hsRow,
                                         const char* __restrict__ ringB, int& off, int ringBytes, bool slot0)
{
    constexpr int KT = kMacKT;
#pragma unroll
    for (int u = 0; u < NT; ++u)
    {
        const double2 h = *reinterpret_cast<const double2*>(hsRow + u * kMacRowBytes);
        const double2 incoming = nxt;
        off -= kMacRowBytes;
        if (off < 0) off += ringBytes;
        nxt = *reinterpret_cast<const double2*>(ringB + off);   // two frames before w[0] of this tap (unused after the last tap)
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
#pragma unroll
            for (int i = 0; i < KT; ++i)
            {
                const double2 x = w[(i + KT - u) % KT];   // logical w[i] at tap u
                acc[i].x = fma(x.x, h.x, fma(-x.y, h.y, acc[i].x));
                acc[i].y = fma(x.x, h.y, fma(x.y, h.x, acc[i].y));
            }
        }
        // slide: logical w[i] <- w[i-1], w[0] <- incoming; the freed physical slot is (KT-1-u)
        w[(KT - 1 - u) % KT] = incoming;
    }
}

// Three-multiplication complex MAC (Gauss): with s = a + b, (a + ib)(c + id) = [c s - b (c + d)] + i [c s + a (d - c)], so
// the three sums  K1 = sum c s,  K2 = sum a (d - c),  K3 = sum b (c + d)  are accumulated with one DFMA each per tap and
// output frame, and re = K1 - K3, im = K1 + K2 is formed once after the last tap: 24 DFMA + 3 DADD per tap and thread
// (d - c, c + d of the H value; s of the spectrum that enters the window) instead of 32 DFMA.  The rounding error bound is
// that of the four-multiplication form times a small constant (|c||a+b| + |b||c+d| against |ac| + |bd|).
struct MacW3 { double a, b, s; };
// NT consecutive taps; rowOff = byte offset of the ring row holding the frame before logical w[0].  The window slides by
// loading that row straight into the slot whose frame has just left the window (no staging registers, no moves): the slot
// is consumed first in tap u (logical i = KT-1) and its new content last in tap u+1 (logical i = 0), which gives the load
// a whole tap of arithmetic to land.  ALIGNED: the ring has a multiple of eight rows and the run starts on a multiple of
// eight, so the eight rows of a block never wrap and their addresses are immediates.
template <int NT, bool ALIGNED>
__device__ __forceinline__ void mac_taps3(double (&k1)[kMacKT], double (&k2)[kMacKT], double (&k3)[kMacKT], MacW3 (&w)[kMacKT],
                                          const char* __restrict__ hsRow, const char* __restrict__ ringB, int& rowOff, int ringBytes)
{
    constexpr int KT = kMacKT;
#pragma unroll
    for (int u = 0; u < NT; ++u)
    {
        const double2 h = *reinterpret_cast<const double2*>(hsRow + u * kMacRowBytes);
        const double dmc = h.y - h.x, cpd = h.x + h.y;
#pragma unroll
        for (int ii = 0; ii < KT; ++ii)
        {
            const int i = KT - 1 - ii;
            const MacW3& x = w[(i + KT - u) % KT];
            k1[i] = fma(h.x, x.s, k1[i]);
            k2[i] = fma(x.a, dmc, k2[i]);
            k3[i] = fma(x.b, cpd, k3[i]);
        }
        double2 v;
        if (ALIGNED) v = *reinterpret_cast<const double2*>(ringB + rowOff - u * kMacRowBytes);
        else
        {
            v = *reinterpret_cast<const double2*>(ringB + rowOff);
            rowOff -= kMacRowBytes;
            if (rowOff < 0) rowOff += ringBytes;
        }
        MacW3& slot = w[(KT - 1 - u) % KT];
        slot.a = v.x;
        slot.b = v.y;
        slot.s = v.x + v.y;
    }
    if (ALIGNED)
    {
        rowOff -= NT * kMacRowBytes;
        if (rowOff < 0) rowOff += ringBytes;
    }
}

template <bool ALIGNED>
__device__ __forceinline__ void mac_run3(double2 (&acc)[kMacKT], const char* __restrict__ hsB, const char* __restrict__ ringB, int off,
                                         int ringBytes, int nq)
{
    constexpr int KT = kMacKT;
    MacW3 w[KT];
    double k1[KT], k2[KT], k3[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i)
    {
        k1[i] = k2[i] = k3[i] = 0.0;
        int o = off + i * kMacRowBytes;
        if (!ALIGNED && o >= ringBytes) o -= ringBytes;
        const double2 v = *reinterpret_cast<const double2*>(ringB + o);
        w[i].a = v.x;
        w[i].b = v.y;
        w[i].s = v.x + v.y;
    }
    int rowOff = off - kMacRowBytes;
    if (rowOff < 0) rowOff += ringBytes;
    const int nFull = nq & ~(KT - 1);
    for (int q0 = 0; q0 < nFull; q0 += KT) mac_taps3<KT, ALIGNED>(k1, k2, k3, w, hsB + q0 * kMacRowBytes, ringB, rowOff, ringBytes);
    const char* hsT = hsB + nFull * kMacRowBytes;
    switch (nq & (KT - 1))
    {
        case 1: mac_taps3<1, ALIGNED>(k1, k2, k3, w, hsT, ringB, rowOff, ringBytes); break;
        case 2: mac_taps3<2, ALIGNED>(k1, k2, k3, w, hsT, ringB, rowOff, ringBytes); break;
        case 3: mac_taps3<3, ALIGNED>(k1, k2, k3, w, hsT, ringB, rowOff, ringBytes); break;
        case 4: mac_taps3<4, ALIGNED>(k1, k2, k3, w, hsT, ringB, rowOff, ringBytes); break;
        case 5: mac_taps3<5, ALIGNED>(k1, k2, k3, w, hsT, ringB, rowOff, ringBytes); break;
        case 6: mac_taps3<6, ALIGNED>(k1, k2, k3, w, hsT, ringB, rowOff, ringBytes); break;
        case 7: mac_taps3<7, ALIGNED>(k1, k2, k3, w, hsT, ringB, rowOff, ringBytes); break;
        default: break;
    }
#pragma unroll
    for (int i = 0; i < KT; ++i) acc[i] = make_double2(k1[i] - k3[i], k1[i] + k2[i]);
}

template <bool PACKED>
__device__ __forceinline__ void mac_run(double2 (&acc)[kMacKT], const char* __restrict__ hsB, const char* __restrict__ ringB, int off,
                                        int ringBytes, int nq, bool slot0)
{
    constexpr int KT = kMacKT;
    double2 w[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i)
    {
        acc[i] = make_double2(0.0, 0.0);
        int o = off + i * kMacRowBytes;
        if (o >= ringBytes) o -= ringBytes;
        w[i] = *reinterpret_cast<const double2*>(ringB + o);   // frame ks + i - qBegin
    }
    off -= kMacRowBytes;
    if (off < 0) off += ringBytes;
    double2 nxt = *reinterpret_cast<const double2*>(ringB + off);   // frame ks - qBegin - 1
    const int nFull = nq & ~(KT - 1);
    for (int q0 = 0; q0 < nFull; q0 += KT) mac_taps<KT, PACKED>(acc, w, nxt, hsB + q0 * kMacRowBytes, ringB, off, ringBytes, slot0);
    const char* hsT = hsB + nFull * kMacRowBytes;
    switch (nq & (KT - 1))
    {
        case 1: mac_taps<1, PACKED>(acc, w, nxt, hsT, ringB, off, ringBytes, slot0); break;
        case 2: mac_taps<2, PACKED>(acc, w, nxt, hsT, ringB, off, ringBytes, slot0); break;
        case 3: mac_taps<3, PACKED>(acc, w, nxt, hsT, ringB, off, ringBytes, slot0); break;
        case 4: mac_taps<4, PACKED>(acc, w, nxt, hsT, ringB, off, ringBytes, slot0); break;
        case 5: mac_taps<5, PACKED>(acc, w, nxt, hsT, ringB, off, ringBytes, slot0); break;
        case 6: mac_taps<6, PACKED>(acc, w, nxt, hsT, ringB, off, ringBytes, slot0); break;
        case 7: mac_taps<7, PACKED>(acc, w, nxt, hsT, ringB, off, ringBytes, slot0); break;
        default: break;
    }
}

#ifndef CPQ_MAC_ALIGNED
#define CPQ_MAC_ALIGNED 1
#endif
#ifndef CPQ_MAC_STAGE1
#define CPQ_MAC_STAGE1 0   // one issuing lane per warp instead of eight: measured slower (MAC 18.6 vs 17.2 ms per step)
#endif
#ifndef CPQ_MAC_GAUSS
#define CPQ_MAC_GAUSS 1
#endif
#ifndef CPQ_MAC_MINBLOCKS
#define CPQ_MAC_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(kMacThreads, CPQ_MAC_MINBLOCKS) mac_kernel(MacArgs a)
{
    extern __shared__ __align__(128) unsigned char mac_smem[];
    const int nq = a.qEnd - a.qBegin;
    const int R = a.ringRows;
    double2* Hs = reinterpret_cast<double2*>(mac_smem);   // [nq][32]
    double2* ring = Hs + nq * kMacBins;                   // [R][32], row of frame f lives in slot f mod R
    uint64_t* bar = reinterpret_cast<uint64_t*>(ring + (size_t) R * kMacBins);
    const int tid = threadIdx.x;
    const int ml = tid & (kMacBins - 1);
    const int g = tid / kMacBins;       // thread group == warp
    const int m0 = blockIdx.x * kMacBins;          // P is a multiple of 32: tiles are always full
    const int seq = blockIdx.z;
    const int kc0 = blockIdx.y * a.framesPerCta;
    const int kc1 = min(a.K, kc0 + a.framesPerCta);
    const int hrow = a.hSeqMod > 0 ? ((a.seqBase + seq) % a.hSeqMod) : seq;
    const double2* __restrict__ X = a.X + ((size_t) seq * a.xRows + a.hist) * a.P + m0;   // frame 0
    const double2* __restrict__ H = a.H + (size_t) hrow * a.hSeqStride + (size_t) a.qBegin * a.P + m0;
    double2* __restrict__ Y = a.Y + (size_t) seq * a.K * a.P + m0 + ml;
    const int ringBytes = R * kMacRowBytes;

    if (tid == 0) mbar_init(bar, 1);

    // Stage rows [f0, f1) of X into the ring (and, with withH, the H tile): frames before 0 are the zero history of a
    // Reset engine (plain stores by all threads), frames in [0, K) one bulk copy each, issued by warp 0.
    auto stage = [&](int f0, int f1, bool withH) {
        const int z1 = min(f1, -a.hist);
        if (f0 < z1)
        {
            const int n = (z1 - f0) * kMacBins;
            for (int i = tid; i < n; i += kMacThreads)
            {
                int s = (f0 + i / kMacBins) % R;
                if (s < 0) s += R;
                ring[s * kMacBins + (i & (kMacBins - 1))] = make_double2(0.0, 0.0);
            }
            fence_proxy_async();
        }
        // bulk copies: 8 rows per warp (lanes 0..7), H rows on lanes 8..15; the expected byte count is posted by thread 0
        // (the transaction count may run negative until then, the phase cannot complete before that arrive)
        const int c0 = max(f0, -a.hist), c1 = min(f1, a.K);
        const int nRows = max(c1 - c0, 0);
        if (tid == 0) mbar_arrive_expect_tx(bar, (unsigned) (nRows + (withH ? nq : 0)) * kMacRowBytes);
#if CPQ_MAC_STAGE1
        // one lane per warp issues the warp's share: rows c0 + [g*per, (g+1)*per), ring slot advanced with a wrap instead of
        // a modulo per row (a divergent multi-lane issue costs an ELECT loop of eight instructions per copy)
        if (ml == 0)
        {
            const int per = (nRows + kMacGroups - 1) / kMacGroups;
            const int i0 = g * per, i1 = min(nRows, i0 + per);
            if (i0 < i1)
            {
                int slot = (c0 + i0) % R;
                if (slot < 0) slot += R;
                const double2* src = X + (int64_t) (c0 + i0) * a.P;
                for (int i = i0; i < i1; ++i)
                {
                    bulk_g2s(ring + slot * kMacBins, src, kMacRowBytes, bar);
                    src += a.P;
                    if (++slot == R) slot = 0;
                }
            }
            if (withH)
            {
                const int perH = (nq + kMacGroups - 1) / kMacGroups;
                const int h1 = min(nq, (g + 1) * perH);
                for (int i = g * perH; i < h1; ++i) bulk_g2s(Hs + i * kMacBins, H + (size_t) i * a.P, kMacRowBytes, bar);
            }
        }
#else
        if (ml < 8)
        {
            for (int i = g * 8 + ml; i < nRows; i += 8 * kMacGroups)
            {
                const int f = c0 + i;
                int slot = f % R;
                if (slot < 0) slot += R;
                bulk_g2s(ring + slot * kMacBins, X + (int64_t) f * a.P, kMacRowBytes, bar);
            }
        }
        else if (withH && ml < 16)
        {
            for (int i = g * 8 + ml - 8; i < nq; i += 8 * kMacGroups) bulk_g2s(Hs + i * kMacBins, H + (size_t) i * a.P, kMacRowBytes, bar);
        }
#endif
    };

    __syncthreads();   // mbarrier initialised
    stage(kc0 - a.qBegin - (nq - 1), kc0 - a.qBegin + kMacSuper, true);
    mbar_wait(bar, 0);
    __syncthreads();   // zero rows written by other threads

    const char* ringB = reinterpret_cast<const char*>(ring) + ml * (int) sizeof(double2);
    const char* hsB = reinterpret_cast<const char*>(Hs) + ml * (int) sizeof(double2);
    const bool packedTile = (m0 == 0);
    const bool alignedRing = CPQ_MAC_ALIGNED && (R % kMacKT) == 0 && (a.qBegin % kMacKT) == 0;   // see mac_taps3
    const bool slot0 = packedTile && ml == 0;
    unsigned phase = 0;
    for (int ks0 = kc0; ks0 < kc1; ks0 += kMacSuper)
    {
        const bool more = ks0 + kMacSuper < kc1;
        // prefetch the rows that only the next super-step needs
        if (more) stage(ks0 - a.qBegin + kMacSuper, ks0 - a.qBegin + 2 * kMacSuper, false);
        const int ks = ks0 + g * kMacKT;
        if (ks < kc1)
        {
            double2 acc[kMacKT];
            int s = (ks - a.qBegin) % R;
            if (s < 0) s += R;
            if (packedTile) mac_run<true>(acc, hsB, ringB, s * kMacRowBytes, ringBytes, nq, slot0);
#if CPQ_MAC_GAUSS
            else if (alignedRing) mac_run3<true>(acc, hsB, ringB, s * kMacRowBytes, ringBytes, nq);
            else mac_run3<false>(acc, hsB, ringB, s * kMacRowBytes, ringBytes, nq);
#else
            else mac_run<false>(acc, hsB, ringB, s * kMacRowBytes, ringBytes, nq, slot0);
#endif
#pragma unroll
            for (int i = 0; i < kMacKT; ++i)
                if (ks + i < kc1) Y[(size_t) (ks + i) * a.P] = acc[i];
        }
        if (more)
        {
            phase ^= 1u;
            mbar_wait(bar, phase);
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------
// The same multiply-accumulate with both operands staged by tensor-map TMA copies (cp.async.bulk.tensor, SASS UTMALDG):
// X is a rank-3 tensor [sequence][frame][bin], a box of 64 frames x 32 bins (32 KB) is ONE copy instead of 64 row copies
// issued through an ELECT loop, frames outside [-hist, K) -- the zero history of a Reset engine -- are the out-of-bounds fill
// of the copy engine (no zero stores, no proxy fence), and the H tile [Q][32] is one copy as well.  The ring is NB blocks of
// 64 rows on full / empty mbarriers: a warp waits only for the block it is about to read, the producer (thread 0) only for
// the block it is about to overwrite -- there is no CTA barrier in the frame loop (the row-copy kernel spent 8 % of its
// samples on that barrier and 10 % waiting for rows it had requested one super-step earlier: here the request for block s + 1
// goes out before block s is touched and block s + 1 is not needed before super-step s + 1).  Used when the tap count is at
// most 65 (one block of history); the row-copy kernel above keeps the uniform-partition extension's longer filters.
// ---------------------------------------------------------------------------------------------
constexpr int kMacTmaNB = 3;                       // history | current | prefetch
constexpr int kMacTmaMaxTaps = kMacSuper + 1;      // nq - 1 <= 64
inline size_t macTmaSmemBytes(int nq, int groups = kMacGroups)
{
    return ((size_t) nq * kMacRowBytes + 127) / 128 * 128 + (size_t) kMacTmaNB * groups * kMacKT * kMacRowBytes + 128;
}

__device__ __forceinline__ void tma_load_3d(void* dstSmem, const void* tmap, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dstSmem)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct alignas(64) MacTensorMap { unsigned char bytes[128]; };   // CUtensorMap (opaque here; encoded on the host)

// GROUPS thread groups of 32 bins x 8 output frames: 8 (blocks of 64 frames, 256 threads) for whole signals, 4 (blocks of 32
// frames, 128 threads) for short calls -- streaming calls of a few callbacks and the dither's time segments, where a layer has
// fewer than 64 frames per call and the 64-frame block would leave half the CTA's warps without work.  Needs nq - 1 <= 8 GROUPS.
template <int GROUPS>
__global__ void __launch_bounds__(kMacBins * GROUPS, GROUPS == kMacGroups ? CPQ_MAC_MINBLOCKS : 4)
mac_tma_kernel(MacArgs a, const __grid_constant__ MacTensorMap tmX, const __grid_constant__ MacTensorMap tmH)
{
    extern __shared__ __align__(128) unsigned char mac_smem[];
    constexpr int NB = kMacTmaNB;
    constexpr int kSuper = GROUPS * kMacKT;
    constexpr int R = NB * kSuper;
    const int nq = a.qEnd - a.qBegin;
    const size_t hBytes = ((size_t) nq * kMacRowBytes + 127) / 128 * 128;
    double2* Hs = reinterpret_cast<double2*>(mac_smem);                      // [nq][32]
    double2* ring = reinterpret_cast<double2*>(mac_smem + hBytes);            // [R][32]: frame f in row (f - base) mod R
    uint64_t* full = reinterpret_cast<uint64_t*>(mac_smem + hBytes + (size_t) R * kMacRowBytes);   // [NB]
    uint64_t* empty = full + NB;                                              // [NB]
    uint64_t* hbar = empty + NB;
    const int tid = threadIdx.x;
    const int ml = tid & (kMacBins - 1);
    const int g = tid / kMacBins;
    const int m0 = blockIdx.x * kMacBins;
    const int seq = blockIdx.z;
    const int kc0 = blockIdx.y * a.framesPerCta;
    const int kc1 = min(a.K, kc0 + a.framesPerCta);
    const int hrow = a.hSeqMod > 0 ? ((a.seqBase + seq) % a.hSeqMod) : (a.seqBase + seq);
    double2* __restrict__ Y = a.Y + (size_t) seq * a.K * a.P + m0 + ml;
    constexpr int ringBytes = R * kMacRowBytes;
    const int base = kc0 - a.qBegin;                  // frame held by ring row 0 of block 0
    const int jmin = nq > 1 ? -1 : 0;                 // first block this CTA ever loads (one block of history)

    if (tid == 0)
    {
        for (int i = 0; i < NB; ++i)
        {
            mbar_init(full + i, 1);
            mbar_init(empty + i, GROUPS);
        }
        mbar_init(hbar, 1);
    }
    __syncthreads();
    // block j = frames base + 64 j .. + 63 -> ring slot j mod NB; the k-th use of a slot completes phase k of its barriers
    auto slotOf = [&](int j) { return ((j - jmin) % NB); };
    auto useOf = [&](int j) { return (j - jmin) / NB; };
    auto loadBlock = [&](int j) {
        const int sl = slotOf(j);
        mbar_arrive_expect_tx(full + sl, (unsigned) kSuper * kMacRowBytes);
        tma_load_3d(ring + (size_t) sl * kSuper * kMacBins, &tmX, 2 * m0, base + kSuper * j + a.hist, seq, full + sl);
    };
    if (tid == 0)
    {
        mbar_arrive_expect_tx(hbar, (unsigned) nq * kMacRowBytes);
        tma_load_3d(Hs, &tmH, 2 * m0, a.qBegin, hrow, hbar);
        for (int j = jmin; j <= 0; ++j) loadBlock(j);
    }
    mbar_wait(hbar, 0);
    if (jmin < 0) mbar_wait(full + slotOf(jmin), 0);

    // ring row of frame f: (f - base) mod R, with block jmin in slot 0: row = (f - base - 64 jmin) mod R
    const int rowShift = -kSuper * jmin;
    const char* ringB = reinterpret_cast<const char*>(ring) + ml * (int) sizeof(double2);
    const char* hsB = reinterpret_cast<const char*>(Hs) + ml * (int) sizeof(double2);
    const bool packedTile = (m0 == 0);
    const bool slot0 = packedTile && ml == 0;
    const int nSteps = (kc1 - kc0 + kSuper - 1) / kSuper;
    for (int s = 0; s < nSteps; ++s)
    {
        if (tid == 0 && s + 1 < nSteps)
        {
            // block s + 1 replaces block s + 1 - NB, which every warp released at the end of super-step s - 1
            const int j = s + 1;
            if (j - NB >= jmin) mbar_wait(empty + slotOf(j), (unsigned) (useOf(j) - 1) & 1u);
            loadBlock(j);
        }
        mbar_wait(full + slotOf(s), (unsigned) useOf(s) & 1u);
        const int ks = kc0 + s * kSuper + g * kMacKT;
        if (ks < kc1)
        {
            double2 acc[kMacKT];
            int r = (ks - a.qBegin - base + rowShift) % R;     // = (ks - kc0 + rowShift) mod R, never negative
            if (packedTile) mac_run<true>(acc, hsB, ringB, r * kMacRowBytes, ringBytes, nq, slot0);
#if CPQ_MAC_GAUSS
            else mac_run3<true>(acc, hsB, ringB, r * kMacRowBytes, ringBytes, nq);   // R and the run start are multiples of eight
#else
            else mac_run<false>(acc, hsB, ringB, r * kMacRowBytes, ringBytes, nq, slot0);
#endif
#pragma unroll
            for (int i = 0; i < kMacKT; ++i)
                if (ks + i < kc1) Y[(size_t) (ks + i) * a.P] = acc[i];
        }
        // this warp is done with the oldest block of the window (block s - 1; with one tap there is no history block: block s)
        __syncwarp();
        if (ml == 0)
        {
            const int jr = nq > 1 ? s - 1 : s;
            if (jr >= jmin) mbar_arrive(empty + slotOf(jr));
        }
    }
}

} // namespace cpq
