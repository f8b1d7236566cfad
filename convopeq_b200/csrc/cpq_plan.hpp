// cpq_plan.hpp -- host-side integer/scalar logic of the hot path (no CUDA in this header).
//
// The reference's real-time machinery (ring buffers, frequency-domain delay line slots, time-sliced tail
// MAC, delay-line cursors) has only *integer* observable consequences for an offline run: which layer
// output sample lands on which output sample, and which callbacks skip a layer.  This header derives
// those from the same rules the reference applies, so that the GPU can compute every layer as one plain
// batched block convolution and then follow a small gather plan.
//
// Reference (paths relative to the reference's src/):
//   layer plan / gains          MKLNonUniformConvolver.cpp:626-684, 738-758, 784-786, 988-994, 1004-1024
//   callback state machine      MKLNonUniformConvolver.cpp:1407-1548 (Add), 1553-1688 (Get, delay line)
//   spectrum filter / tilt      MKLNonUniformConvolver.cpp:336-443, 1060-1097
//   SVF design                  eqprocessor/EQProcessor.Coefficients.cpp:84-130, 431-618
//   total-gain ramp             DspNumericPolicy.h:319-421, eqprocessor/EQProcessor.Processing.cpp:1262-1274
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cctype>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/cpq.h"

namespace cpq
{

constexpr double kPi = 3.14159265358979323846;

inline int nextPow2(int n)
{
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}
inline int ilog2(int n)
{
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}
template <typename T>
inline T clampv(T lo, T hi, T v) { return v < lo ? lo : (hi < v ? hi : v); }

struct LayerPlan
{
    int partSize = 0, fftSize = 0, bins = 0;       // P, N = 2P, M = P + 1
    int numPartsIR = 0, numParts = 0;              // Q, nextPow2(Q)
    int partsPerCallback = 0, callbacksPerCycle = 0;
    int outputDelaySamples = 0;                    // sum of previous layers' IR lengths
    int irOffset = 0, irLen = 0;
    bool immediate = false;
    double gain = 1.0;
};

struct ConvPlan
{
    int numLayers = 0;
    int blockSize = 0;   // SetImpulse's blockSize (the host's block rounded up to a power of two, LoaderThread.cpp:230)
    int callSize = 0;    // samples per Add/Get call = the host's block (preferredCallSize, :239-245); == blockSize for power-of-two hosts
    int irLen = 0;
    int tailMode = 1;
    bool tailEnabled = true;
    double tailStartSec = 0.085, strength01 = 0.5;
    bool hasSpec = false;
    cpq_filter_spec spec {};
    LayerPlan layers[CPQ_MAX_LAYERS];

    bool sameGeometry(const ConvPlan& o) const
    {
        if (numLayers != o.numLayers || blockSize != o.blockSize || callSize != o.callSize) return false;
        for (int l = 0; l < numLayers; ++l)
        {
            const LayerPlan &a = layers[l], &b = o.layers[l];
            if (a.partSize != b.partSize || a.numPartsIR != b.numPartsIR || a.partsPerCallback != b.partsPerCallback ||
                a.outputDelaySamples != b.outputDelaySamples || a.gain != b.gain)
                return false;
        }
        return true;
    }
};

inline void filterSpecDefault(cpq_filter_spec* s)
{
    s->sample_rate = 48000.0;
    s->hc_mode = 1;
    s->lc_mode = 0;
    s->tail_mode = 1;
    s->tail_enabled = 1;
    s->tail_start_seconds = 0.085;
    s->tail_strength = 1.0;
    s->tail_l1l2_multiplier = 8;
    s->reserved_ = 0;
}

// SetImpulse's parameter table and layer plan.  Returns false for the inputs SetImpulse rejects.
// uniformExtension: BASELINE config 1's "uniform partitioned convolution" beyond what the reference can express -- its L0 is
// capped at 32 partitions (kL0MaxParts, :744-747), so 65,536 taps never run as 128 uniform 512-sample partitions there.  With
// the flag the whole IR is one immediate layer of ceil(irLen / P0) partitions (no tails, no layer gains): an extension,
// checked against linear convolution instead of the reference.
inline bool makeConvPlan(int irLen, int blockSize, const cpq_filter_spec* fs, ConvPlan& out, bool uniformExtension = false)
{
    if (irLen <= 0 || blockSize <= 0) return false;
    out = ConvPlan {};
    out.blockSize = blockSize;
    out.callSize = blockSize;
    out.irLen = irLen;
    out.hasSpec = fs != nullptr;
    if (fs) out.spec = *fs;

    const int tailMode = fs ? clampv(0, 2, (int) fs->tail_mode) : 1;
    const bool tailEnabled = (tailMode != 2) && (fs ? fs->tail_enabled != 0 : true);
    const double srTail = fs ? fs->sample_rate : 48000.0;
    double tailStart = fs ? clampv(0.01, 0.80, fs->tail_start_seconds) : 0.085;
    const double userStrength = fs ? clampv(0.0, 2.0, fs->tail_strength) : 1.0;
    double tailStrength = userStrength;
    int mult = fs ? clampv(2, 16, (int) fs->tail_l1l2_multiplier) : 8;
    double g1 = 1.0, g2 = 1.0;
    const double s01 = clampv(0.0, 1.0, userStrength * 0.5);
    if (!tailEnabled) { g1 = 0.0; g2 = 0.0; }
    else if (tailMode == 0)
    {
        tailStart = clampv(0.01, 0.80, std::max(tailStart, 0.055));
        mult = clampv(2, 16, std::max(mult, 6));
        tailStrength = clampv(0.0, 2.0, userStrength);
        g1 = clampv(0.0, 2.0, tailStrength * (0.95 - 0.25 * s01));
        g2 = clampv(0.0, 2.0, tailStrength * (0.80 - 0.45 * s01));
    }
    else if (tailMode == 1)
    {
        tailStart = clampv(0.01, 0.80, std::max(tailStart, 0.12));
        tailStrength = clampv(0.0, 2.0, std::max(tailStrength, 1.25));
        mult = clampv(2, 16, std::max(mult, 8));
        g1 = clampv(0.0, 2.0, tailStrength * (1.05 + 0.20 * s01));
        g2 = clampv(0.0, 2.0, tailStrength * (0.82 + 0.12 * s01));
    }
    else { g1 = 0.0; g2 = 0.0; }
    out.tailMode = tailMode;
    out.tailEnabled = tailEnabled;
    out.tailStartSec = tailStart;
    out.strength01 = s01;

    const int l0Part = nextPow2(std::max(blockSize, 64));
    const long long l1PartLL = (long long) l0Part * mult;
    const long long l2PartLL = l1PartLL * mult;
    const int l0MaxLen = 32 * l0Part;
    const int l0ByTail = (int) std::llround(tailStart * srTail);
    const int l0Target = clampv(l0Part, l0MaxLen, l0ByTail);
    const int l0Len = uniformExtension ? irLen : std::min(irLen, tailEnabled ? l0Target : l0MaxLen);
    const long long l1Cap = 64LL * l1PartLL;
    const int l1Len = (tailEnabled && !uniformExtension) ? (int) std::max(0LL, std::min((long long) irLen - l0Len, l1Cap)) : 0;
    const int l2Len = (tailEnabled && !uniformExtension) ? std::max(0, irLen - l0Len - l1Len) : 0;

    const int offs[3] = { 0, l0Len, l0Len + l1Len };
    const int lens[3] = { l0Len, l1Len, l2Len };
    const long long parts[3] = { l0Part, l1PartLL, l2PartLL };
    const double gains[3] = { 1.0, g1, g2 };

    int prevTotal = 0;
    for (int li = 0; li < 3; ++li)
    {
        if (lens[li] <= 0) continue;
        if (parts[li] > (1LL << 28)) return false;
        LayerPlan& l = out.layers[out.numLayers];
        l.partSize = (int) parts[li];
        l.fftSize = l.partSize * 2;
        l.bins = l.partSize + 1;
        l.immediate = (li == 0);
        l.numPartsIR = (lens[li] + l.partSize - 1) / l.partSize;
        l.numParts = nextPow2(l.numPartsIR);
        l.irOffset = offs[li];
        l.irLen = lens[li];
        // the reference indexes m_tailLayerGain by *active* layer index (Get(), :1626-1628)
        l.gain = tailEnabled ? gains[out.numLayers] : 0.0;
        if (!l.immediate)
        {
            const int bs = std::max(blockSize, 1);
            const int blocksPerPart = (l.partSize + bs - 1) / bs;
            int ppc = std::max(1, (l.numPartsIR + blocksPerPart - 1) / blocksPerPart);
            ppc = std::min(ppc, l.numPartsIR);
            l.partsPerCallback = ppc;
            l.callbacksPerCycle = (l.numPartsIR + ppc - 1) / ppc;
        }
        l.outputDelaySamples = prevTotal;
        prevTotal += lens[li];
        ++out.numLayers;
    }
    if (out.numLayers > 0) out.layers[0].gain = 1.0;
    return out.numLayers > 0;
}

// Real per-bin gain applied to every partition spectrum of layer `li` (applySpectrumFilter + tilt).
inline void spectrumGain(const ConvPlan& plan, int li, std::vector<double>& gain)
{
    const LayerPlan& l = plan.layers[li];
    gain.assign((size_t) l.bins, 1.0);
    if (plan.hasSpec)
    {
        const cpq_filter_spec& spec = plan.spec;
        const double fs = spec.sample_rate;
        const double nyquist = fs * 0.5;
        const double hcStart = (fs <= 48000.0) ? 18000.0 : 22000.0;
        const double hcEnd = nyquist;
        const double lcEnd = (spec.lc_mode == 1) ? 6.0 : 8.0;
        const double lcStart = (spec.lc_mode == 1) ? 15.0 : 18.0;
        const int N = l.fftSize, halfN = N / 2, cs = l.bins;
        {
            const int kStart = (int) std::round(hcStart * N / fs);
            const int kEnd = std::min(halfN, (int) std::round(hcEnd * N / fs));
            for (int k = 0; k < cs; ++k)
            {
                if (k <= kStart) continue;
                if (k <= kEnd)
                {
                    const double denom = (double) (kEnd - kStart);
                    const double x = (double) (k - kStart) / denom;
                    switch (spec.hc_mode)
                    {
                        case 0: gain[(size_t) k] = 1.0 / std::sqrt(1.0 + std::pow(x, 8.0)); break;
                        case 1: gain[(size_t) k] = 0.5 * (1.0 + std::cos(kPi * x)); break;
                        case 2: gain[(size_t) k] = std::exp(-4.60517 * x * x); break;
                        default: break;
                    }
                }
            }
        }
        {
            const int kEnd = (int) std::round(lcEnd * N / fs);
            const int kStart = (int) std::round(lcStart * N / fs);
            for (int k = 0; k < cs; ++k)
            {
                if (k <= kEnd) gain[(size_t) k] = 0.0;
                else if (k < kStart)
                {
                    const double denom = (double) std::max(1, kStart - kEnd);
                    const double x = (double) (k - kEnd) / denom;
                    gain[(size_t) k] *= 0.5 * (1.0 - std::cos(kPi * x));
                }
            }
        }
    }
}

// Air-absorption tilt for layers >= 1 in tail mode 0; returns false when no tilt applies.
inline bool tiltGain(const ConvPlan& plan, int li, std::vector<double>& tilt)
{
    if (!(plan.tailEnabled && plan.tailMode == 0) || li < 1) return false;
    const LayerPlan& l = plan.layers[li];
    const double startNorm = clampv(0.65, 1.55, plan.tailStartSec / 0.085);
    const double dampingBase = (0.35 + 1.10 * plan.strength01) * startNorm;
    const double w = (li == 1) ? 1.0 : 1.6;
    const double dc = dampingBase * w;
    const double denom = (double) std::max(1, l.bins - 1);
    tilt.resize((size_t) l.bins);
    for (int k = 0; k < l.bins; ++k)
    {
        const double fn = (double) k / denom;
        tilt[(size_t) k] = std::exp(-dc * fn * fn);
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Gather plan: run the reference's per-callback integer state machine for `nCallbacks` callbacks of
// exactly blockSize samples (Add then Get).
//
//   L0   : ring of completed L0 blocks; per callback (ringSrc, count): out[0..count) = y0[ringSrc ...],
//          rest zero.  With blockSize == partSize this is the identity (ringSrc = c*B, count = B).
//   L>=1 : the delay line is the concatenation of *completed* blocks (a block whose distribution cycle
//          is restarted before it finished is never written); per callback either -1 (skipped) or the
//          stream position read.  blockOfStream maps stream block j to the frame index k it came from.
// ------------------------------------------------------------------------------------------------
struct GatherPlan
{
    int64_t nCallbacks = 0;
    std::vector<int64_t> l0Src;                    // [nCallbacks] or empty when identity
    std::vector<int32_t> l0Count;                  // [nCallbacks]
    bool l0Identity = true;
    std::vector<int64_t> tailSrc[CPQ_MAX_LAYERS];  // [nCallbacks], -1 = skip   (index = layer)
    std::vector<int32_t> blockOfStream[CPQ_MAX_LAYERS];  // stream block -> frame k; empty when identity
    bool blockIdentity[CPQ_MAX_LAYERS] = { true, true, true };
    int64_t firstOutput[CPQ_MAX_LAYERS] = { -1, -1, -1 };
    int64_t skipped[CPQ_MAX_LAYERS] = { 0, 0, 0 };
    int64_t framesNeeded[CPQ_MAX_LAYERS] = { 0, 0, 0 };  // frames k < framesNeeded must be computed
};

inline void simulateCallbacks(const ConvPlan& plan, int64_t nCallbacks, GatherPlan& g)
{
    g = GatherPlan {};
    g.nCallbacks = nCallbacks;
    const int B = plan.callSize > 0 ? plan.callSize : plan.blockSize;

    // ---- L0 ring (ringWrite/ringRead; capacity never binds when Get follows every Add) ----
    {
        const LayerPlan& l = plan.layers[0];
        const int P = l.partSize;
        g.framesNeeded[0] = (nCallbacks * (int64_t) B) / P;
        if (P == B) g.l0Identity = true;
        else
        {
            g.l0Identity = false;
            g.l0Src.resize((size_t) nCallbacks);
            g.l0Count.resize((size_t) nCallbacks);
            int64_t written = 0, read = 0;
            int inputPos = 0;
            for (int64_t c = 0; c < nCallbacks; ++c)
            {
                int consumed = 0;
                while (consumed < B)
                {
                    const int fill = std::min(B - consumed, P - inputPos);
                    inputPos += fill;
                    consumed += fill;
                    if (inputPos >= P) { inputPos = 0; written += P; }
                }
                const int64_t avail = written - read;
                const int cnt = (int) std::min<int64_t>(B, avail);
                g.l0Src[(size_t) c] = read;
                g.l0Count[(size_t) c] = cnt;
                read += cnt;
            }
        }
        g.firstOutput[0] = 0;
    }

    // ---- tails ----
    for (int li = 1; li < plan.numLayers; ++li)
    {
        const LayerPlan& l = plan.layers[li];
        const int P = l.partSize;
        g.tailSrc[li].assign((size_t) nCallbacks, -1);
        std::vector<int32_t> completed;
        int inputPos = 0, nextPart = 0;
        bool distributing = false;
        int32_t frameInFlight = -1, framesPushed = 0;
        uint64_t w = 0, r = 0;
        bool spoke = false;
        for (int64_t c = 0; c < nCallbacks; ++c)
        {
            // Add
            int consumed = 0;
            while (consumed < B)
            {
                const int fill = std::min(B - consumed, P - inputPos);
                inputPos += fill;
                consumed += fill;
                if (inputPos >= P)
                {
                    inputPos = 0;
                    frameInFlight = framesPushed++;
                    nextPart = 0;
                    distributing = true;
                }
            }
            if (distributing)
            {
                const int endPart = std::min(nextPart + l.partsPerCallback, l.numPartsIR);
                nextPart = endPart;
                if (nextPart >= l.numPartsIR)
                {
                    completed.push_back(frameInFlight);
                    w += (uint64_t) P;
                    distributing = false;
                    nextPart = 0;
                }
            }
            // Get -> delayLineReadAdd
            const uint64_t maxRead = (w >= (uint64_t) l.outputDelaySamples) ? w - (uint64_t) l.outputDelaySamples : 0;
            const uint64_t start = std::max(r, maxRead);
            if (start + (uint64_t) B > w)
            {
                if (spoke) ++g.skipped[li];
                continue;
            }
            g.tailSrc[li][(size_t) c] = (int64_t) start;
            r = start + (uint64_t) B;
            if (!spoke) { spoke = true; g.firstOutput[li] = c * (int64_t) B; }
        }
        bool ident = true;
        for (size_t j = 0; j < completed.size(); ++j)
            if (completed[j] != (int32_t) j) { ident = false; break; }
        g.blockIdentity[li] = ident;
        if (!ident) g.blockOfStream[li] = completed;
        g.framesNeeded[li] = completed.empty() ? 0 : (int64_t) (*std::max_element(completed.begin(), completed.end())) + 1;
    }
}

// ------------------------------------------------------------------------------------------------
// EQ design (float parameters clamped, then promoted) and the per-band scan constants.
// ------------------------------------------------------------------------------------------------
inline void bypassCoeffs(cpq_svf_coeffs* c) { c->a1 = 1.0; c->a2 = 0.0; c->a3 = 0.0; c->m0 = 1.0; c->m1 = 0.0; c->m2 = 0.0; }

inline bool designBand(int type, float freq, float gainDb, float q, double sr, cpq_svf_coeffs* out)
{
    if (type < 0 || type > 4) return false;
    if (sr <= 0.0) { bypassCoeffs(out); return true; }
    const float nyq = static_cast<float>(sr * 0.5);
    const float maxFreq = std::min(20000.0f, nyq * 0.95f);
    freq = clampv(20.0f, maxFreq, freq);
    q = clampv(0.01f, 20.0f, q);
    gainDb = clampv(-48.0f, 48.0f, gainDb);
    const double f = (double) freq, gdb = (double) gainDb, Q = (double) q;
    double A = 1.0, g = 0.0, k = 0.0;
    const double t = std::tan(kPi * f / sr);
    switch (type)
    {
        case 0: A = std::pow(10.0, gdb / 40.0); g = t / std::sqrt(A); k = 1.0 / Q; break;
        case 1: A = std::pow(10.0, gdb / 40.0); g = t; k = 1.0 / (Q * A); break;
        case 2: A = std::pow(10.0, gdb / 40.0); g = t * std::sqrt(A); k = 1.0 / Q; break;
        default: g = t; k = 1.0 / Q; break;
    }
    if (!std::isfinite(g) || !std::isfinite(k)) { bypassCoeffs(out); return true; }
    const double den = 1.0 + g * (g + k);
    if (std::fabs(den) < 1.0e-15) { bypassCoeffs(out); return true; }
    out->a1 = 1.0 / den;
    out->a2 = g * out->a1;
    out->a3 = g * out->a2;
    switch (type)
    {
        case 0: out->m0 = 1.0; out->m1 = k * (A - 1.0); out->m2 = A * A - 1.0; break;
        case 1: out->m0 = 1.0; out->m1 = (A - 1.0 / A) / Q; out->m2 = 0.0; break;
        case 2: out->m0 = A * A; out->m1 = k * (1.0 - A) * A; out->m2 = 1.0 - A * A; break;
        case 3: out->m0 = 0.0; out->m1 = 0.0; out->m2 = 1.0; break;
        default: out->m0 = 1.0; out->m1 = -k; out->m2 = -1.0; break;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// IR preparation (host, one-time): IRConverter::computeScaleFactor (IRConverter.cpp:13-196) with
// IRAnalyzer::estimateMaxFrequencyResponseGain (IRAnalyzer.cpp:63-155).
// ------------------------------------------------------------------------------------------------
// IR preparation between the (optional) resampler and SetImpulse (LoaderThread::doLoadStep,
// convolver/ConvolverProcessor.LoaderThread.cpp:588-637): per channel a 1 Hz UltraHighRateDCBlocker
// (UltraHighRateDCBlocker.h:78-188), the asymmetric Tukey window around the peak (applyAsymmetricTukey,
// ConvolverProcessor.ResampleAndFallback.cpp:111-196), then the copy into targetLength samples
// (computeTargetIRLength, ConvolverProcessor.StateAndUI.cpp:942-957) with a linear fade-out over the last 2 % of the copied
// samples (256 .. 80 ms).  Resampling itself is r8brain-free-src (third party, vendored by the reference) and is not restated.
// ------------------------------------------------------------------------------------------------
inline void irDcBlock(double* data, int n, double sr, double cutoffHz)
{
    double alpha[2] = { 1.0e-6, 1.0e-6 };
    if (std::isfinite(sr) && sr > 0.0 && std::isfinite(cutoffHz) && cutoffHz > 0.0)
    {
        const double ratios[2] = { 1.0 - 0.1, 1.0 + 0.1 };
        for (int i = 0; i < 2; ++i)
        {
            const double omega = 2.0 * kPi * (cutoffHz * ratios[i]) / sr;
            double a = -std::expm1(-omega);
            if (!std::isfinite(a) || a <= 0.0 || a >= 1.0) a = 1.0e-6;
            alpha[i] = a;
        }
    }
    double s0 = 0.0, s1 = 0.0;
    for (int i = 0; i < n; ++i)
    {
        double x = data[i];
        s0 = std::fma(alpha[0], x - s0, s0);   // the reference's build contracts state + alpha * (x - state)
        x -= s0;
        s1 = std::fma(alpha[1], x - s1, s1);
        x -= s1;
        data[i] = x;
    }
}

inline void irAsymmetricTukey(double* data, int n)
{
    if (!data || n <= 0) return;
    int peak = 0;
    for (int i = 1; i < n; ++i)
        if (std::fabs(data[peak]) < std::fabs(data[i])) peak = i;   // std::max_element: first of equal maxima
    const double alphaPre = 0.05;
    double alphaPost = 0.05 + 0.033 * (std::log2((double) n) - 10.0);
    alphaPost = std::max(0.05, std::min(0.25, alphaPost));
    if (peak > 0)
    {
        const int len = (int) std::floor(peak * alphaPre);
        const double scale = kPi / (peak * alphaPre);
        for (int i = 0; i < len; ++i) data[i] *= 0.5 * (1.0 + std::cos(scale * (double) i + -kPi));
    }
    const double dist = (double) (n - 1 - peak);
    if (dist > 1.0e-9)
    {
        const int start = peak + (int) std::ceil(dist * (1.0 - alphaPost));
        const int len = n - start;
        if (len > 0)
        {
            const double scale = (kPi / alphaPost) / dist;
            const double offset = (kPi / alphaPost) * (((double) start - (double) peak) / dist - (1.0 - alphaPost));
            for (int i = 0; i < len; ++i) data[start + i] *= 0.5 * (1.0 + std::cos(scale * (double) i + offset));
        }
    }
}

inline int irTargetLength(double sr, double targetSeconds)
{
    int target = (int) (sr * (double) (float) targetSeconds);   // targetIRLengthSec is a float
    target = std::min(target, 2097152);                          // MAX_IR_LATENCY, ConvolverProcessor.h:198
    return std::max(target, 1);
}

// in[len] at the device rate -> out[targetLength] (zero padded); returns targetLength.  out must hold irTargetLength() samples.
inline int irPrepare(const double* in, int len, double sr, double targetSeconds, double* out)
{
    const int target = irTargetLength(sr, targetSeconds);
    std::vector<double> w(in, in + std::max(len, 0));
    if (sr > 0.0 && len > 0) irDcBlock(w.data(), len, sr, 1.0);
    if (len > 0) irAsymmetricTukey(w.data(), len);
    std::fill(out, out + target, 0.0);
    const int copy = std::min(target, std::max(len, 0));
    int fade = (int) std::round((double) copy * 0.02);
    const int maxFade = std::max(256, (int) std::round(sr * 0.080));
    fade = std::max(256, std::min(maxFade, fade));
    fade = std::max(0, std::min(fade, copy - 1));
    std::copy(w.begin(), w.begin() + copy, out);
    if (fade > 0)
    {
        // juce::AudioBuffer::applyGainRamp(start, n, 1, 0): gain = 1 + i * (0 - 1) / n
        const double inc = (0.0 - 1.0) / (double) fade;
        double g = 1.0;
        for (int i = 0; i < fade; ++i)
        {
            out[copy - fade + i] *= g;
            g += inc;
        }
    }
    return target;
}

// ------------------------------------------------------------------------------------------------
inline double irFreqPeakGain(const double* const* ch, int nch, int len)
{
    if (len <= 0 || nch <= 0) return 1.0;
    const int copyLen = std::min(len, 65536);   // kMaxAnalysisWindow
    int n = 1;
    while (n < copyLen) n <<= 1;
    if (n < 2) return 1.0;
    // Tukey window, alpha = 0.5, defined over the FFT length; coherent gain = its mean over the copied samples
    const double span = 0.5 * (double) (n - 1), taper = 0.5 * span;
    std::vector<double> win((size_t) n);
    for (int i = 0; i < n; ++i)
    {
        const double t = (double) i;
        if (t < taper) win[(size_t) i] = 0.5 * (1.0 + std::cos(2.0 * kPi * t / span - kPi));
        else if (t > (double) (n - 1) - taper) win[(size_t) i] = 0.5 * (1.0 + std::cos(2.0 * kPi * (t - ((double) (n - 1) - taper)) / span));
        else win[(size_t) i] = 1.0;
    }
    double mean = 0.0;
    for (int i = 0; i < copyLen; ++i) mean += win[(size_t) i];
    mean /= (double) copyLen;
    if (mean < 1e-18) return 1.0;
    int log2n = 0;
    while ((1 << log2n) < n) ++log2n;
    std::vector<std::complex<double>> tw((size_t) n / 2), z((size_t) n);
    for (int k = 0; k < n / 2; ++k) tw[(size_t) k] = std::polar(1.0, -2.0 * kPi * (double) k / (double) n);
    std::vector<double> mag((size_t) n / 2 + 1);
    double best = 0.0;
    for (int c = 0; c < nch; ++c)
    {
        if (!ch[c]) continue;
        for (int i = 0; i < n; ++i)
        {
            unsigned r = 0;
            for (int b = 0; b < log2n; ++b) r |= ((unsigned) i >> b & 1u) << (log2n - 1 - b);
            z[r] = i < copyLen ? ch[c][i] * win[(size_t) i] : 0.0;
        }
        for (int half = 1; half < n; half <<= 1)
            for (int i = 0; i < n; i += 2 * half)
                for (int j = 0; j < half; ++j)
                {
                    const std::complex<double> t = tw[(size_t) j * (size_t) (n / (2 * half))] * z[(size_t) (i + j + half)];
                    const std::complex<double> u = z[(size_t) (i + j)];
                    z[(size_t) (i + j)] = u + t;
                    z[(size_t) (i + j + half)] = u - t;
                }
        const int nb = n / 2;
        for (int b = 0; b <= nb; ++b)
        {
            mag[(size_t) b] = (b == 0 || b == nb) ? std::fabs(z[(size_t) b].real()) : std::abs(z[(size_t) b]);
            best = std::max(best, mag[(size_t) b]);
        }
        for (int b = 1; b < nb - 1; ++b)   // three-point log-parabolic refinement of local maxima
        {
            const double ym = mag[(size_t) b - 1], y0 = mag[(size_t) b], yp = mag[(size_t) b + 1];
            if (!(y0 > ym && y0 > yp && y0 > 1e-18 && ym > 1e-18 && yp > 1e-18)) continue;
            const double lm = std::log(ym), l0 = std::log(y0), lp = std::log(yp), den = lm - 2.0 * l0 + lp;
            if (std::fabs(den) > 1e-18) best = std::max(best, y0 * std::exp(-(0.5 * (lm - lp) / den) * (l0 - lm)));
        }
    }
    best /= mean;
    return best > 1e-18 ? best : 1.0;
}

// In-place radix-2 complex FFT (host, prepare-time only); sign -1 forward, +1 backward (unscaled).
inline void hostFft(std::vector<std::complex<double>>& z, const std::vector<std::complex<double>>& tw, int sign)
{
    const size_t n = z.size();
    for (size_t i = 1, j = 0; i < n; ++i)
    {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(z[i], z[j]);
    }
    for (size_t half = 1; half < n; half <<= 1)
    {
        const size_t step = n / (2 * half);
        for (size_t i = 0; i < n; i += 2 * half)
            for (size_t j = 0; j < half; ++j)
            {
                const std::complex<double> w = sign < 0 ? tw[j * step] : std::conj(tw[j * step]);
                const std::complex<double> t = w * z[i + j + half], u = z[i + j];
                z[i + j] = u + t;
                z[i + j + half] = u - t;
            }
    }
}

// convertToMinimumPhase, convolver/ConvolverProcessor.ResampleAndFallback.cpp:333-460: homomorphic (real-cepstrum folding)
// minimum-phase reconstruction on an FFT of nextPow2(4 len) points.  Returns false where the reference returns an empty
// buffer (FFT above 2^23 points, non-finite result), in which case the loader keeps the linear-phase IR.
inline bool irMinimumPhase(const double* ir, int len, double* out)
{
    if (!ir || !out || len <= 0) return false;
    size_t n = 1;
    while (n < (size_t) len * 4) n <<= 1;
    if (n > 8388608) return false;   // MAX_MINPHASE_FFT_SIZE
    std::vector<std::complex<double>> tw(n / 2), z(n);
    for (size_t k = 0; k < n / 2; ++k) tw[k] = std::polar(1.0, -2.0 * kPi * (double) k / (double) n);
    for (size_t i = 0; i < n; ++i) z[i] = i < (size_t) len ? ir[i] : 0.0;
    hostFft(z, tw, -1);
    for (auto& v : z) v = std::log(std::max(std::abs(v), 1.0e-300));   // log magnitude, zero phase
    hostFft(z, tw, +1);
    const double inv = 1.0 / (double) n;
    const size_t half = n / 2;
    for (size_t i = 0; i < n; ++i)   // fold the real cepstrum onto its causal part
    {
        const double c = z[i].real() * inv;
        z[i] = (i == 0 || i == half) ? c : (i < half ? 2.0 * c : 0.0);
    }
    hostFft(z, tw, -1);
    for (auto& v : z)
    {
        v = std::exp(std::complex<double>(clampv(-50.0, 50.0, v.real()), clampv(-50.0, 50.0, v.imag())));
        if (!std::isfinite(v.real()) || !std::isfinite(v.imag())) return false;
    }
    hostFft(z, tw, +1);
    for (int i = 0; i < len; ++i)
    {
        double v = z[(size_t) i].real() * inv;
        if (!std::isfinite(v)) return false;
        if (std::fabs(v) < 1.0e-18) v = 0.0;
        out[i] = v;
    }
    return true;
}

// convertToMixedPhaseFallback ("Phase 1"), convolver/ConvolverProcessor.MixedPhase.cpp:721-865: the magnitude of the linear-phase
// IR with a phase that follows the minimum-phase IR below loHz, a pure delay (the linear IR's peak) above hiHz and a raised-cosine
// blend between -- one complex FFT of nextPow2(len) points per channel.  The reference's first choice, an optimised all-pass
// cascade with a disk cache (:68-719), is a design tool and is not restated.  Returns false where the reference returns empty.
inline bool irMixedPhaseFallback(const double* lin, const double* minp, int len, double sr, double loHz, double hiHz, double* out)
{
    if (!lin || !minp || !out || len <= 0 || !(sr > 0.0) || !(hiHz > loHz)) return false;
    size_t n = 1;
    while (n < (size_t) len) n <<= 1;
    if (n > 8388608) return false;   // MAX_MIXED_FFT_SIZE
    const size_t half = n / 2, nc = half + 1;
    std::vector<std::complex<double>> tw(std::max<size_t>(n / 2, 1)), zl(n), zm(n);
    for (size_t k = 0; k < tw.size(); ++k) tw[k] = std::polar(1.0, -2.0 * kPi * (double) k / (double) n);
    int peakDelay = 0;
    double maxVal = 0.0;
    for (int i = 0; i < len; ++i)
        if (std::fabs(lin[i]) > maxVal)
        {
            maxVal = std::fabs(lin[i]);
            peakDelay = i;
        }
    for (size_t i = 0; i < n; ++i)
    {
        zl[i] = i < (size_t) len ? lin[i] : 0.0;
        zm[i] = i < (size_t) len ? minp[i] : 0.0;
    }
    hostFft(zl, tw, -1);
    hostFft(zm, tw, -1);
    std::vector<double> dphi(nc);
    const double invSpan = 1.0 / (hiHz - loHz);
    for (size_t k = 0; k < nc; ++k)
    {
        const double freq = ((double) k * sr) / (double) n;
        double wMin = 1.0;
        if (freq >= hiHz) wMin = 0.0;
        else if (freq > loHz) wMin = 0.5 * (1.0 + std::cos(kPi * ((freq - loHz) * invSpan)));
        const double wLin = 1.0 - wMin;
        const double phiLin = -(2.0 * kPi * (double) k / (double) n) * (double) peakDelay;
        const double phiMin = std::atan2(zm[k].imag(), zm[k].real());
        dphi[k] = (wLin * phiLin + wMin * phiMin) - phiLin;
    }
    // unwrapPhaseRadians, ConvolverProcessor.Internal.h:33-46, as written: the step is taken against the already corrected neighbour
    double correction = 0.0;
    for (size_t i = 1; i < nc; ++i)
    {
        const double delta = dphi[i] - dphi[i - 1];
        if (delta > kPi) correction -= 2.0 * kPi;
        else if (delta < -kPi) correction += 2.0 * kPi;
        dphi[i] += correction;
    }
    for (size_t k = 0; k < n; ++k)
    {
        const double d = k <= half ? dphi[k] : -dphi[n - k];
        const double re = zl[k].real(), im = zl[k].imag(), c = std::cos(d), sn = std::sin(d);
        zl[k] = std::complex<double>(re * c - im * sn, re * sn + im * c);
    }
    hostFft(zl, tw, +1);
    const double inv = 1.0 / (double) n;
    for (int i = 0; i < len; ++i)
    {
        const double v = zl[(size_t) i].real() * inv;
        out[i] = std::fabs(v) < 1.0e-18 ? 0.0 : v;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// IR file decode.  The reference reads IR files through JUCE (juce::AudioFormatManager::registerBasicFormats ->
// WavAudioFormat, convolver/ConvolverProcessor.LoaderThread.cpp:439-486; JUCE is an external dependency that is not part of
// the reference tree): the reader delivers 32-bit floats -- integer PCM is left-justified to 32 bits and multiplied by
// 1.0f / 0x7fffffff (= 2^-31 in float), IEEE samples are narrowed to float -- which convertFloatToDoubleHighQuality
// (InputBitDepthTransform.h:102-121) widens and passes through the input transform (NaN / |v| < 1e-20 -> 0, clamp to
// [-1, 1]).  Restated for RIFF/WAVE PCM 8/16/24/32, IEEE float 32/64 and WAVE_FORMAT_EXTENSIBLE with those sub-formats.
// ------------------------------------------------------------------------------------------------
struct WavData
{
    int channels = 0;
    double sampleRate = 0.0;
    int bitsPerSample = 0, isFloat = 0;
    std::vector<std::vector<double>> ch;   // [channels][frames]
};

// applyHighQuality64BitTransform with gain 1 (InputBitDepthTransform.h:86-100): the four-wide body lets +-Inf through to the clamp,
// the scalar remainder (n % 4 samples) zeroes it -- the same rule as input_kernel
inline void irInputTransform(double* d, int n)
{
    const int vecEnd = n / 4 * 4;
    for (int i = 0; i < n; ++i)
    {
        double v = d[i];
        if (v != v || std::fabs(v) < 1.0e-20 || (i >= vecEnd && std::isinf(v))) v = 0.0;
        d[i] = std::min(std::max(v, -1.0), 1.0);
    }
}

inline bool irDecodeWav(const uint8_t* b, size_t n, WavData& out, std::string& err)
{
    auto u16 = [&](size_t o) { return (uint32_t) b[o] | ((uint32_t) b[o + 1] << 8); };
    auto u32 = [&](size_t o) { return (uint32_t) b[o] | ((uint32_t) b[o + 1] << 8) | ((uint32_t) b[o + 2] << 16) | ((uint32_t) b[o + 3] << 24); };
    if (n < 12 || std::memcmp(b, "RIFF", 4) != 0 || std::memcmp(b + 8, "WAVE", 4) != 0)
    {
        err = "not a RIFF/WAVE file";
        return false;
    }
    int tag = 0, channels = 0, bits = 0, align = 0;
    double rate = 0.0;
    size_t dataOff = 0, dataLen = 0;
    for (size_t o = 12; o + 8 <= n;)
    {
        const size_t len = u32(o + 4), body = o + 8;
        if (std::memcmp(b + o, "fmt ", 4) == 0 && len >= 16 && body + 16 <= n)
        {
            tag = (int) u16(body);
            channels = (int) u16(body + 2);
            rate = (double) u32(body + 4);
            align = (int) u16(body + 12);
            bits = (int) u16(body + 14);
            if (tag == 0xFFFE && len >= 26 && body + 26 <= n) tag = (int) u16(body + 24);   // WAVE_FORMAT_EXTENSIBLE: first word of the sub-format GUID
        }
        else if (std::memcmp(b + o, "data", 4) == 0)
        {
            dataOff = body;
            dataLen = std::min(len, n - body);
            break;
        }
        o = body + len + (len & 1);
    }
    const bool isFloat = tag == 3;
    if ((tag != 1 && tag != 3) || channels <= 0 || rate <= 0.0 || dataOff == 0 ||
        !(isFloat ? (bits == 32 || bits == 64) : (bits == 8 || bits == 16 || bits == 24 || bits == 32)))
    {
        err = "unsupported WAVE format (PCM 8/16/24/32 and IEEE float 32/64 are read)";
        return false;
    }
    const int bytes = bits / 8;
    if (align < bytes * channels) align = bytes * channels;
    const size_t frames = dataLen / (size_t) align;
    out.channels = channels;
    out.sampleRate = rate;
    out.bitsPerSample = bits;
    out.isFloat = isFloat ? 1 : 0;
    out.ch.assign((size_t) channels, std::vector<double>(frames));
    const float kFixed = 1.0f / (float) 0x7fffffff;
    for (size_t f = 0; f < frames; ++f)
        for (int c = 0; c < channels; ++c)
        {
            const uint8_t* p = b + dataOff + f * (size_t) align + (size_t) c * bytes;
            float v;
            if (isFloat && bits == 32)
            {
                uint32_t w = (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24);
                std::memcpy(&v, &w, 4);
            }
            else if (isFloat)
            {
                uint64_t w = 0;
                for (int i = 7; i >= 0; --i) w = (w << 8) | p[i];
                double d;
                std::memcpy(&d, &w, 8);
                v = (float) d;
            }
            else
            {
                int32_t s;
                if (bits == 8) s = (int32_t) (((uint32_t) p[0] - 128u) << 24);
                else if (bits == 16) s = (int32_t) (((uint32_t) p[0] << 16) | ((uint32_t) p[1] << 24));
                else if (bits == 24) s = (int32_t) (((uint32_t) p[0] << 8) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 24));
                else s = (int32_t) ((uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24));
                v = (float) s * kFixed;
            }
            out.ch[(size_t) c][f] = (double) v;
        }
    for (auto& v : out.ch) irInputTransform(v.data(), (int) v.size());
    return true;
}

// doTrimStep's trailing-silence trim (LoaderThread.cpp:496-547): the length up to the last sample above 1e-15 in either of the
// first two channels, at least 1
inline int irTrimTrailingSilence(const double* ch0, const double* ch1, int n)
{
    int len = 0;
    for (int j = n - 1; j >= 0; --j)
        if (std::fabs(ch0[j]) > 1.0e-15 || (ch1 && std::fabs(ch1[j]) > 1.0e-15))
        {
            len = j + 1;
            break;
        }
    return std::max(1, len);
}

inline void irPeakAndRms(const double* const* ch, int nch, int len, double scale, double& peak, double& rms)
{
    double e = 0.0;
    peak = 0.0;
    for (int c = 0; c < nch; ++c)
        for (int i = 0; i < len; ++i)
        {
            const double v = ch[c][i] * scale;
            peak = std::max(peak, std::fabs(v));
            e += v * v;
        }
    rms = (nch * len > 0) ? std::sqrt(e / (double) (nch * len)) : 0.0;
}

inline void irScaleFactor(const double* const* ch, int nch, int len, const double* const* cur, int curCh, int curLen, double curScale,
                          cpq_ir_scale* out)
{
    out->scale_factor = 1.0;
    out->has_scale_factor = 0;
    out->additional_attenuation_db = 0.0f;
    double scale = 1.0;
    if (len > 0 && nch > 0)
    {
        double maxEnergy = 0.0;   // computeEnergyScale: loudest channel to unit energy, then -6 dB
        for (int c = 0; c < nch; ++c)
        {
            double e = 0.0;
            for (int i = 0; i < len; ++i) e += ch[c][i] * ch[c][i];
            if (std::isfinite(e) && e > 1.0e-18) maxEnergy = std::max(maxEnergy, e);
        }
        if (maxEnergy > 1.0e-18 && std::isfinite(maxEnergy)) scale = (1.0 / std::sqrt(maxEnergy)) * 0.5011872336272722;
    }
    if (!(scale > 0.0) || !std::isfinite(scale)) return;
    out->has_scale_factor = 1;
    double result = scale, peak = 0.0, rms = 0.0;
    if (len > 0 && nch > 0) irPeakAndRms(ch, nch, len, 1.0, peak, rms);
    const double fgain = (len > 0 && nch > 0) ? irFreqPeakGain(ch, nch, len) : 1.0;   // of the unscaled IR, like analyzeIR
    double attDb = 0.0;
    if (peak * scale > 0.5)   // kMaxEffectivePeak
    {
        const double k = 0.5 / (peak * scale);
        result *= k;
        scale *= k;
        attDb += -20.0 * std::log10(k);
    }
    if (rms * scale > 0.25)   // kMaxEffectiveRms, judged after the peak clamp
    {
        const double k = 0.25 / (rms * scale);
        result *= k;
        attDb += -20.0 * std::log10(k);
    }
    if (fgain > 1.41)         // kMaxEffectiveFreqResponse (+3 dB)
    {
        const double k = 1.41 / fgain;
        result *= k;
        attDb += -20.0 * std::log10(k);
    }
    out->additional_attenuation_db = (float) attDb;
    if (cur && curCh > 0 && curLen > 0)   // jump protection against the IR that is playing
    {
        double cp, cr, np_, nr;
        irPeakAndRms(cur, curCh, curLen, curScale, cp, cr);
        irPeakAndRms(ch, nch, len, result, np_, nr);
        const bool peakJump = cp > 1.0e-9 && np_ > cp * 4.0 && np_ > 0.5, rmsJump = cr > 1.0e-9 && nr > cr * 4.0 && nr > 0.25;
        if (peakJump || rmsJump)
        {
            double kp = std::numeric_limits<double>::infinity(), kr = kp;
            if (np_ > 1.0e-12 && cp > 1.0e-12) kp = cp * 4.0 / np_;
            if (nr > 1.0e-12 && cr > 1.0e-12) kr = cr * 4.0 / nr;
            const double k = std::min(kp, kr);
            if (std::isfinite(k) && k > 0.0 && k < 1.0) result *= k;
        }
    }
    out->scale_factor = result;
}

// ------------------------------------------------------------------------------------------------
// EQ preset text (EqualizerAPO / AutoEq "ParametricEq.txt"): EQProcessor::loadFromTextFile,
// eqprocessor/EQProcessor.Core.cpp:300-495.  Host-only; fills the same parameter fields the reference's setters would.
// ------------------------------------------------------------------------------------------------
namespace preset
{
inline bool startsWithNoCase(const std::string& s, const char* p)
{
    size_t i = 0;
    for (; p[i]; ++i)
        if (i >= s.size() || std::tolower((unsigned char) s[i]) != std::tolower((unsigned char) p[i])) return false;
    return true;
}
inline bool equalsNoCase(const std::string& s, const char* p)
{
    size_t i = 0;
    for (; p[i]; ++i)
        if (i >= s.size() || std::tolower((unsigned char) s[i]) != std::tolower((unsigned char) p[i])) return false;
    return i == s.size();
}
// juce::String::getFloatValue: the leading number of the token (0 when there is none), narrowed to float
inline float floatValue(const std::string& s) { return (float) std::strtod(s.c_str(), nullptr); }
}   // namespace preset

inline int parseEqPreset(const char* text, cpq_eq_band_params bands[CPQ_NUM_BANDS], float* totalGainDb)
{
    using namespace preset;
    for (int i = 0; i < CPQ_NUM_BANDS; ++i)   // :305-310
    {
        bands[i].enabled = 0;
        bands[i].channel_mode = 0;
        bands[i].gain_db = 0.0f;
    }
    int filterIndex = 0, ignored = 0, mode = 0;
    const std::string all(text ? text : "");
    size_t pos = 0;
    while (pos <= all.size())
    {
        size_t eol = all.find('\n', pos);
        if (eol == std::string::npos) eol = all.size();
        std::string line = all.substr(pos, eol - pos);
        pos = eol + 1;
        line = line.substr(0, line.find('#'));     // upToFirstOccurrenceOf("#"), then ";" (:320-322)
        line = line.substr(0, line.find(';'));
        std::vector<std::string> tok;
        for (size_t i = 0; i < line.size();)
        {
            while (i < line.size() && std::isspace((unsigned char) line[i])) ++i;
            size_t j = i;
            while (j < line.size() && !std::isspace((unsigned char) line[j])) ++j;
            if (j > i) tok.push_back(line.substr(i, j - i));
            i = j;
        }
        if (tok.empty()) continue;
        if (startsWithNoCase(tok[0], "Preamp"))   // :339-350: the first later token that contains a digit, '-' or '.'
        {
            for (size_t i = 1; i < tok.size(); ++i)
                if (tok[i].find_first_of("0123456789-.") != std::string::npos)
                {
                    *totalGainDb = clampv(-48.0f, 48.0f, floatValue(tok[i]));   // setTotalGain clamps (Parameters.cpp:106)
                    break;
                }
        }
        else if (startsWithNoCase(tok[0], "Channel"))   // :351-374
        {
            bool hasL = false, hasR = false;
            for (std::string t : tok)
            {
                if (startsWithNoCase(t, "Channel")) t = t.substr(7);
                std::string u;
                for (char c : t)
                    if (c != ':' && c != ',') u.push_back(c);
                if (equalsNoCase(u, "L") || equalsNoCase(u, "Left")) hasL = true;
                else if (equalsNoCase(u, "R") || equalsNoCase(u, "Right")) hasR = true;
            }
            mode = (hasL && hasR) ? 0 : (hasL ? 1 : (hasR ? 2 : 0));
        }
        else if (startsWithNoCase(tok[0], "Filter"))   // :375-486
        {
            if (filterIndex >= CPQ_NUM_BANDS) { ++ignored; continue; }
            cpq_eq_band_params& b = bands[filterIndex];
            bool enabled = true, typeFound = false, qFound = false;
            float freq = 0.0f, gain = 0.0f, q = 0.707f;
            for (size_t i = 1; i < tok.size(); ++i)
            {
                const std::string& t = tok[i];
                if (equalsNoCase(t, "ON")) { enabled = true; continue; }
                if (equalsNoCase(t, "OFF")) { enabled = false; continue; }
                if (!typeFound)
                {
                    int type = -1;
                    if (equalsNoCase(t, "LSC") || equalsNoCase(t, "LowShelf")) type = 0;
                    else if (equalsNoCase(t, "PK") || equalsNoCase(t, "Peaking")) type = 1;
                    else if (equalsNoCase(t, "HSC") || equalsNoCase(t, "HighShelf")) type = 2;
                    else if (equalsNoCase(t, "LP") || equalsNoCase(t, "LowPass")) type = 3;
                    else if (equalsNoCase(t, "HP") || equalsNoCase(t, "HighPass")) type = 4;
                    if (type >= 0) { b.type = type; typeFound = true; continue; }
                }
                if (i + 1 < tok.size())
                {
                    if (equalsNoCase(t, "Fc")) freq = floatValue(tok[i + 1]);
                    else if (equalsNoCase(t, "Gain")) gain = floatValue(tok[i + 1]);
                    else if (equalsNoCase(t, "Q")) { q = floatValue(tok[i + 1]); qFound = true; }
                }
            }
            if (!typeFound) b.type = 1;
            b.enabled = enabled ? 1 : 0;
            if (freq > 0.0f) b.frequency = freq;
            b.gain_db = gain;
            b.q = (qFound && q > 0.0f) ? q : 0.707f;
            b.channel_mode = mode;
            ++filterIndex;
        }
    }
    return ignored;
}

inline double dbToGain(float db)
{
    const double d = (double) db;
    return d > -100.0 ? std::pow(10.0, d * 0.05) : 0.0;
}

inline double equalPowerSin(double x)
{
    const double t = x * (kPi * 0.5);
    const double t2 = t * t;
    return t * (1.0 + t2 * (-1.0 / 6.0 + t2 * (1.0 / 120.0 + t2 * (-1.0 / 5040.0 + t2 * (1.0 / 362880.0)))));
}

// Total-gain LinearRamp evaluated per callback: (start, increment) pairs.
struct GainEvent { int64_t atCallback; double target; };

// The LinearRamp's state between callbacks (DspNumericPolicy.h:319-421): what a stream carries from one call to the next
struct GainRampState
{
    double current = 1.0, target = 1.0, step = 0.0;
    int remaining = 0;      // samples the ramp still has to go; 0 = settled at `target`
    bool live = false;      // false: settled at the parameter set's total gain (no ramp has been started since Reset)
};

inline void gainRampTable(double initial, int totalSteps, int blockSize, int64_t nCallbacks,
                          std::vector<GainEvent> events, std::vector<double>& startInc /* [nCallbacks][2] */, GainRampState* carry = nullptr)
{
    std::stable_sort(events.begin(), events.end(), [](const GainEvent& a, const GainEvent& b) { return a.atCallback < b.atCallback; });
    double current = initial, target = initial, step = 0.0, wanted = initial;
    int remaining = 0;
    if (carry && carry->live)   // a stream continues: the ramp goes on from where the previous call left it
    {
        current = carry->current;
        target = wanted = carry->target;
        step = carry->step;
        remaining = carry->remaining;
    }
    size_t ev = 0;
    startInc.resize((size_t) nCallbacks * 2);
    for (int64_t c = 0; c < nCallbacks; ++c)
    {
        while (ev < events.size() && events[ev].atCallback <= c) wanted = events[ev++].target;
        if (std::fabs(target - wanted) > 1e-6 && wanted != target)
        {
            target = wanted;
            const int steps = remaining > 0 ? remaining : totalSteps;
            step = (target - current) / (double) steps;
            remaining = steps;
        }
        const double start = current;
        if (remaining > 0)
        {
            if (blockSize >= remaining) { current = target; remaining = 0; }
            else { current += step * (double) blockSize; remaining -= blockSize; }
        }
        startInc[(size_t) c * 2] = start;
        startInc[(size_t) c * 2 + 1] = (current - start) / (double) blockSize;
    }
    if (carry)
    {
        // events beyond this call would be lost: the caller schedules relative to the call that follows
        carry->current = current;
        carry->target = target;
        carry->step = step;
        carry->remaining = remaining;
        carry->live = true;
    }
}

// PsychoacousticDither coefficient selection (PsychoacousticDither.h:192-275)
inline void ditherCoeffs(double sampleRate, int bitDepth, double* out12)
{
    static const double table[6][3][12] = {
        { { 2.93, -5.06, 6.97, -7.66, 7.11, -5.63, 3.96, -2.18, 0.80, -0.24, 0.10, -0.04 },
          { 2.49, -4.30, 5.92, -6.51, 6.05, -4.79, 3.37, -1.86, 0.68, -0.20, 0.08, -0.03 },
          { 2.04, -3.52, 4.85, -5.34, 4.95, -3.92, 2.76, -1.52, 0.56, -0.17, 0.07, -0.03 } },
        { { 2.85, -4.92, 6.78, -7.45, 6.92, -5.48, 3.85, -2.12, 0.78, -0.23, 0.09, -0.04 },
          { 2.42, -4.18, 5.75, -6.32, 5.87, -4.65, 3.27, -1.80, 0.66, -0.20, 0.08, -0.03 },
          { 1.98, -3.42, 4.71, -5.18, 4.81, -3.81, 2.68, -1.47, 0.54, -0.16, 0.06, -0.03 } },
        { { 3.28, -5.66, 7.80, -8.57, 7.96, -6.30, 4.43, -2.44, 0.90, -0.27, 0.11, -0.05 },
          { 2.78, -4.80, 6.61, -7.26, 6.75, -5.34, 3.75, -2.07, 0.76, -0.23, 0.09, -0.04 },
          { 2.28, -3.94, 5.42, -5.95, 5.53, -4.38, 3.08, -1.69, 0.62, -0.19, 0.07, -0.03 } },
        { { 3.71, -6.40, 8.82, -9.69, 9.00, -7.12, 5.01, -2.76, 1.02, -0.31, 0.12, -0.05 },
          { 3.15, -5.44, 7.50, -8.24, 7.65, -6.05, 4.25, -2.34, 0.86, -0.26, 0.10, -0.04 },
          { 2.58, -4.46, 6.15, -6.75, 6.27, -4.96, 3.48, -1.92, 0.70, -0.21, 0.08, -0.03 } },
        { { 4.12, -7.10, 9.78, -10.75, 9.98, -7.89, 5.55, -3.06, 1.13, -0.34, 0.14, -0.06 },
          { 3.49, -6.03, 8.31, -9.13, 8.47, -6.70, 4.71, -2.59, 0.95, -0.29, 0.11, -0.05 },
          { 2.86, -4.94, 6.81, -7.48, 6.94, -5.49, 3.86, -2.12, 0.78, -0.23, 0.09, -0.04 } },
        { { 4.48, -7.73, 10.64, -11.70, 10.86, -8.59, 6.04, -3.33, 1.23, -0.37, 0.15, -0.06 },
          { 3.80, -6.56, 9.04, -9.93, 9.22, -7.29, 5.13, -2.82, 1.04, -0.31, 0.12, -0.05 },
          { 3.11, -5.37, 7.41, -8.13, 7.55, -5.97, 4.20, -2.31, 0.85, -0.26, 0.10, -0.04 } },
    };
    int srBand;
    if (sampleRate < 46050.0) srBand = 0;
    else if (sampleRate < 72000.0) srBand = 1;
    else if (sampleRate < 144000.0) srBand = 2;
    else if (sampleRate < 264600.0) srBand = 3;
    else if (sampleRate < 529200.0) srBand = 4;
    else srBand = 5;
    const int bp = bitDepth <= 16 ? 0 : (bitDepth <= 24 ? 1 : 2);
    for (int i = 0; i < 12; ++i) out12[i] = table[srBand][bp][i];
}

} // namespace cpq
