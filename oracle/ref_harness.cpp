// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI harness around the reference's *own, unmodified*
// hot-path classes, compiled in place from $CONVOPEQ_REF/src (never copied into this repo).
// It exists so tests/ and bench.py's cpu_baseline / --impl reference leg can run the real
// reference algorithm: convo::MKLNonUniformConvolver (src/MKLNonUniformConvolver.{h,cpp}) and
// EQProcessor::process(block, params, cache) (src/eqprocessor/EQProcessor.Processing.cpp:1019).
// The product path (convopeq_b200/, include/) never links or loads this file.
//
// EQProcessor.Core.cpp needs the real JUCE and is not linked; the four members it would provide
// (ctor, dtor, getTotalGain, retireBandNodeDeferred) are defined here as the minimum needed to
// construct the object, and prepareToPlay()'s effect on the fields that process() reads
// (EQProcessor.Core.cpp:679-826: maxInternalBlockSize, smoothTotalGain, bypassFadeGain,
// filterState, totalGainTarget) is applied directly.
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <optional>
#include <thread>
#include <vector>
#include <functional>
#include <string>
#include <type_traits>
#include <utility>
#include <limits>
#include <new>
#include <bit>
#include <immintrin.h>

#define private public
#define protected public
#include "PsychoacousticDither.h"   // with ref_shim/mkl_vsl.h standing in for oneMKL VSL (injected uniform streams)
#include "MKLNonUniformConvolver.h"
#include "eqprocessor/EQProcessor.h"
#include "OutputFilter.h"
#undef private
#undef protected
#include "UltraHighRateDCBlocker.h"
#include "IRAnalyzer.h"
#include "SimplePeakLimiter.h"
#include "InputBitDepthTransform.h"

// ---- members normally provided by EQProcessor.Core.cpp -----------------------------------
EQProcessor::EQProcessor()
{
    for (auto& b : bandNodeBits) b.store(0, std::memory_order_relaxed);
}
EQProcessor::~EQProcessor() {}
float EQProcessor::getTotalGain() const { return totalGainDbTarget.load(); }
bool EQProcessor::retireBandNodeDeferred(BandNode* node) noexcept
{
    delete node;
    return true;
}

// ---- injected-uniform VSL streams (ref_shim/mkl_vsl.h) -----------------------------------
std::atomic<uint64_t> convo::PsychoacousticDither::instanceSeedCounterStorage_ { 0 };
std::atomic<uint64_t>& convo::PsychoacousticDither::instanceSeedCounter() noexcept { return instanceSeedCounterStorage_; }
namespace
{
struct VslInject { const double* values = nullptr; long count = 0; };
thread_local VslInject g_vslInject[8];
thread_local int g_vslCreated = 0;
thread_local bool g_vslFail = false;   // vslNewStream reports failure: the reference falls back to its own xorshift64* generator
}
static void cpqref_vsl_begin() { g_vslCreated = 0; for (auto& v : g_vslInject) v = VslInject {}; }
static void cpqref_vsl_inject(int channel, const double* values, long count) { g_vslInject[channel] = VslInject { values, count }; }
int vslNewStream(VSLStreamStatePtr* stream, int, unsigned int)
{
    if (g_vslFail)
    {
        *stream = nullptr;
        return -1;
    }
    auto* s = new cpqref_vsl_stream();
    s->index = g_vslCreated++;
    if (s->index < 8) { s->values = g_vslInject[s->index].values; s->count = g_vslInject[s->index].count; }
    *stream = s;
    return VSL_STATUS_OK;
}
int vslDeleteStream(VSLStreamStatePtr* stream) { delete *stream; *stream = nullptr; return VSL_STATUS_OK; }
int vdRngUniform(int, VSLStreamStatePtr s, MKL_INT n, double* r, double, double)
{
    for (MKL_INT i = 0; i < n; ++i) r[i] = (s->pos < s->count) ? s->values[s->pos++] : 0.5;
    return VSL_STATUS_OK;
}

extern "C" {

struct cpqref_filter_spec
{
    double sample_rate;
    int hc_mode;   // 0 Sharp, 1 Natural, 2 Soft
    int lc_mode;   // 0 Natural, 1 Soft
    int tail_mode; // 0 air absorption, 1 layer tail contouring, 2 bypass
    int tail_enabled;
    double tail_start_seconds;
    double tail_strength;
    int tail_l1l2_multiplier;
};

struct cpqref_eq_band
{
    float frequency, gain_db, q;
    int enabled, type, channel_mode;
};

// ------------------------------------------------------------------ convolver
void* cpqref_nuc_create(void) { return new (std::nothrow) convo::MKLNonUniformConvolver(); }
void cpqref_nuc_destroy(void* h) { delete static_cast<convo::MKLNonUniformConvolver*>(h); }

int cpqref_nuc_set_impulse(void* h, const double* ir, int len, int block, double scale, int direct_head,
                           const cpqref_filter_spec* fs)
{
    auto* c = static_cast<convo::MKLNonUniformConvolver*>(h);
    if (!fs) return c->SetImpulse(ir, len, block, scale, direct_head != 0, nullptr) ? 1 : 0;
    convo::FilterSpec s;
    s.sampleRate = fs->sample_rate;
    s.hcMode = static_cast<convo::HCMode>(fs->hc_mode);
    s.lcMode = static_cast<convo::LCMode>(fs->lc_mode);
    s.tailMode = fs->tail_mode;
    s.tailEnabled = fs->tail_enabled != 0;
    s.tailStartSeconds = fs->tail_start_seconds;
    s.tailStrength = fs->tail_strength;
    s.tailL1L2Multiplier = fs->tail_l1l2_multiplier;
    return c->SetImpulse(ir, len, block, scale, direct_head != 0, &s) ? 1 : 0;
}

// The reference call pattern (StereoConvolver::process, ConvolverProcessor.Runtime.cpp:1159-1184):
// Add(in, n); got = Get(out, n); zero-fill the shortfall.
void cpqref_nuc_process(void* h, const double* in, double* out, long total, int call)
{
    auto* c = static_cast<convo::MKLNonUniformConvolver*>(h);
    juce::ScopedNoDenormals nd;
    for (long pos = 0; pos < total; pos += call)
    {
        const int n = (int) std::min<long>(call, total - pos);
        c->Add(in ? in + pos : nullptr, n);
        const int got = c->Get(out + pos, n);
        if (got < n) std::memset(out + pos + std::max(got, 0), 0, sizeof(double) * (size_t) (n - std::max(got, 0)));
    }
}

void cpqref_nuc_reset(void* h) { static_cast<convo::MKLNonUniformConvolver*>(h)->Reset(); }
int cpqref_nuc_latency(void* h) { return static_cast<convo::MKLNonUniformConvolver*>(h)->getLatency(); }

// layout[l] = {partSize, numPartsIR, numParts, partsPerCallback, outputDelaySamples, delayLineCapacity, isImmediate, fftSize}
int cpqref_nuc_layout(void* h, int* layout /*[3][8]*/, double* gains /*[3]*/)
{
    auto* c = static_cast<convo::MKLNonUniformConvolver*>(h);
    for (int l = 0; l < c->m_numActiveLayers; ++l)
    {
        const auto& L = c->m_layers[l];
        int* o = layout + l * 8;
        o[0] = L.partSize; o[1] = L.numPartsIR; o[2] = L.numParts; o[3] = L.partsPerCallback;
        o[4] = L.outputDelaySamples; o[5] = L.delayLineCapacity; o[6] = L.isImmediate ? 1 : 0; o[7] = L.fftSize;
    }
    for (int l = 0; l < 3; ++l) gains[l] = c->m_tailLayerGain[l];
    return c->m_numActiveLayers;
}

// Copy out one layer's stored partition spectrum (reference order = reversed partitions), SoA.
int cpqref_nuc_ir_spectrum(void* h, int layer, int part, double* re, double* im)
{
    auto* c = static_cast<convo::MKLNonUniformConvolver*>(h);
    if (layer < 0 || layer >= c->m_numActiveLayers) return 0;
    const auto& L = c->m_layers[layer];
    if (part < 0 || part >= L.numParts) return 0;
    std::memcpy(re, L.irFreqReal + (size_t) part * L.complexSize, sizeof(double) * (size_t) L.complexSize);
    std::memcpy(im, L.irFreqImag + (size_t) part * L.complexSize, sizeof(double) * (size_t) L.complexSize);
    return L.complexSize;
}

// ------------------------------------------------------------------ EQ
struct RefEq
{
    EQProcessor proc;
    convo::EQParameters params;
    std::unique_ptr<EQCoeffCache> cache;
    double sr = 48000.0;
    int maxBlock = 512;
};

void* cpqref_eq_create(double sr, int max_block, float total_gain_db)
{
    auto* e = new (std::nothrow) RefEq();
    if (!e) return nullptr;
    e->sr = sr;
    e->maxBlock = max_block;
    auto& p = e->proc;
    // prepareToPlay (EQProcessor.Core.cpp:679-826), the part process() depends on:
    p.currentSampleRate.store(sr);
    p.maxInternalBlockSize = max_block;
    p.smoothTotalGain.totalSteps = convo::LinearRamp::computeTotalSteps(sr, EQProcessor::SMOOTHING_TIME_SEC);
    p.bypassFadeGain.totalSteps = convo::LinearRamp::computeTotalSteps(sr, EQProcessor::BYPASS_FADE_TIME_SEC);
    p.totalGainDbTarget.store(total_gain_db);
    const double lin = juce::Decibels::decibelsToGain<double>(static_cast<double>(total_gain_db));
    p.totalGainTarget.store(lin);
    p.smoothTotalGain.current = p.smoothTotalGain.target = lin;
    p.smoothTotalGain.step = 0.0;
    p.smoothTotalGain.remaining = 0;
    p.bypassFadeGain.current = p.bypassFadeGain.target = 1.0;
    p.bypassFadeGain.step = 0.0;
    p.bypassFadeGain.remaining = 0;
    std::memset(p.filterState.data(), 0, sizeof(p.filterState));
    const int channelRequired = juce::nextPowerOfTwo(max_block) * EQProcessor::MAX_CHANNELS;
    p.parallelInputBuffer = convo::makeAlignedArray<double>((size_t) channelRequired);
    p.parallelWorkBuffer = convo::makeAlignedArray<double>((size_t) channelRequired);
    p.parallelAccumBuffer = convo::makeAlignedArray<double>((size_t) channelRequired);
    p.parallelBufferCapacity = channelRequired;
    p.msWorkBuffer = convo::makeAlignedArray<double>((size_t) channelRequired);   // prepareToPlay: Mid/Side scratch of the node path
    // AGC block-rate coefficients (EQProcessor.Core.cpp:744-792)
    p.agcAttackCoeffTable = convo::makeAlignedArray<double>((size_t) max_block + 1);
    p.agcReleaseCoeffTable = convo::makeAlignedArray<double>((size_t) max_block + 1);
    p.agcSmoothCoeffTable = convo::makeAlignedArray<double>((size_t) max_block + 1);
    p.agcCoeffTableCapacity = max_block + 1;
    p.agcAttackCoeff.store(std::exp(-1.0 / (sr * EQProcessor::AGC_ATTACK_TIME_SEC)));
    p.agcReleaseCoeff.store(std::exp(-1.0 / (sr * EQProcessor::AGC_RELEASE_TIME_SEC)));
    p.agcSmoothCoeff.store(std::exp(-1.0 / (sr * EQProcessor::AGC_SMOOTH_TIME_SEC)));
    for (int i = 0; i <= max_block; ++i)
    {
        const double n = static_cast<double>(i);
        p.agcAttackCoeffTable[i] = 1.0 - std::exp(-n / (sr * EQProcessor::AGC_ATTACK_TIME_SEC));
        p.agcReleaseCoeffTable[i] = 1.0 - std::exp(-n / (sr * EQProcessor::AGC_RELEASE_TIME_SEC));
        p.agcSmoothCoeffTable[i] = 1.0 - std::exp(-n / (sr * EQProcessor::AGC_SMOOTH_TIME_SEC));
    }
    p.rtAgcCurrentGainShadow.store(1.0);
    p.rtAgcEnvInputShadow.store(0.0);
    p.rtAgcEnvOutputShadow.store(0.0);
    return e;
}

void cpqref_eq_destroy(void* h)
{
    auto* e = static_cast<RefEq*>(h);
    if (!e) return;
    delete e->proc.exchangeCurrentState(nullptr);
    for (int i = 0; i < 20; ++i) delete e->proc.exchangeBandNode(i, nullptr);
    delete e;
}

int cpqref_eq_set_params(void* h, const cpqref_eq_band* bands /*[20]*/, float saturation, int structure, int agc)
{
    auto* e = static_cast<RefEq*>(h);
    for (int i = 0; i < 20; ++i)
    {
        auto& b = e->params.bands[(size_t) i];
        b.frequency = bands[i].frequency;
        b.gain = bands[i].gain_db;
        b.q = bands[i].q;
        b.enabled = bands[i].enabled != 0;
        b.type = bands[i].type;
        b.channelMode = bands[i].channel_mode;
    }
    e->params.nonlinearSaturation = saturation;
    e->params.filterStructure = structure;
    e->params.agcEnabled = agc != 0;
    e->cache.reset(EQProcessor::createCoeffCache(e->params, e->sr, e->maxBlock, 1));
    // The node path (process(block), which process(block, params, cache) falls back to for Mid/Side bands,
    // Processing.cpp:1037-1044) reads the published EQState and the per-band BandNodes (createBandNode,
    // EQProcessor.Coefficients.cpp:27-58, incl. its 0-dB skip): publish both the way setBand*/prepareToPlay would.
    auto* st = new EQProcessor::EQState();
    for (int i = 0; i < 20; ++i)
    {
        st->bands[(size_t) i].frequency = bands[i].frequency;
        st->bands[(size_t) i].gain = bands[i].gain_db;
        st->bands[(size_t) i].q = bands[i].q;
        st->bands[(size_t) i].enabled = bands[i].enabled != 0;
        st->bandTypes[(size_t) i] = static_cast<EQBandType>(bands[i].type);
        st->bandChannelModes[(size_t) i] = static_cast<EQChannelMode>(bands[i].channel_mode);
    }
    st->agcEnabled = agc != 0;
    st->nonlinearSaturation = saturation;
    st->filterStructure = structure;
    delete e->proc.exchangeCurrentState(st);
    for (int i = 0; i < 20; ++i) delete e->proc.exchangeBandNode(i, e->proc.createBandNode(i, *st));
    e->proc.rtActiveStructureShadow.store(static_cast<EQProcessor::FilterStructure>(structure));
    return e->cache ? 1 : 0;
}

// setTotalGain (EQProcessor.Parameters.cpp:109 -> storeTotalGainDb): starts the 50 ms ramp at the next process().
void cpqref_eq_set_total_gain(void* h, float db)
{
    auto* e = static_cast<RefEq*>(h);
    e->proc.totalGainDbTarget.store(db);
    e->proc.totalGainTarget.store(juce::Decibels::decibelsToGain<double>(static_cast<double>(db)));
}

void cpqref_eq_process(void* h, double* L, double* R, long total, int block)
{
    auto* e = static_cast<RefEq*>(h);
    for (long pos = 0; pos < total; pos += block)
    {
        const int n = (int) std::min<long>(block, total - pos);
        double* chans[2] = { L + pos, R ? R + pos : nullptr };
        juce::dsp::AudioBlock<double> blk(chans, R ? 2u : 1u, (size_t) n);
        e->proc.process(blk, e->params, e->cache.get());
    }
}

void cpqref_eq_get_state(void* h, double* out /*[2][20][2]*/)
{
    auto* e = static_cast<RefEq*>(h);
    for (int ch = 0; ch < 2; ++ch)
        for (int b = 0; b < 20; ++b)
            for (int k = 0; k < 2; ++k) out[(ch * 20 + b) * 2 + k] = e->proc.filterState[(size_t) ch][(size_t) b][(size_t) k];
}

void cpqref_eq_design(int type, float f, float g, float q, double sr, double* out /*a1,a2,a3,m0,m1,m2*/)
{
    const EQCoeffsSVF c = EQProcessor::calcSVFCoeffs(static_cast<EQBandType>(type), f, g, q, sr);
    out[0] = c.a1; out[1] = c.a2; out[2] = c.a3; out[3] = c.m0; out[4] = c.m1; out[5] = c.m2;
}

// DSPCore::processDouble's ConvolverThenEQ chain per callback (DSPCoreDouble.cpp:386-414,465-469,655-663):
// StereoConvolver::process per channel -> ConvolverProcessor wet scrub + equalPowerSin(1) gain ->
// EQProcessor::process(block, params, cache) -> makeup gain -> kOutputHeadroom.  Used as the CPU baseline.
void cpqref_chain_process(void* nucL, void* nucR, void* eq, double* L, double* R, long total, int block,
                          int outer, double makeup, int epilogue)
{
    auto* cl = static_cast<convo::MKLNonUniformConvolver*>(nucL);
    auto* cr = static_cast<convo::MKLNonUniformConvolver*>(nucR);
    auto* e = static_cast<RefEq*>(eq);
    juce::ScopedNoDenormals nd;
    const double t = 1.0 * (juce::MathConstants<double>::pi * 0.5);
    const double t2 = t * t;
    const double wetG = t * (1.0 + t2 * (-1.0 / 6.0 + t2 * (1.0 / 120.0 + t2 * (-1.0 / 5040.0 + t2 * (1.0 / 362880.0)))));
    for (long pos = 0; pos < total; pos += block)
    {
        const int n = (int) std::min<long>(block, total - pos);
        double* ch[2] = { L + pos, R ? R + pos : nullptr };
        convo::MKLNonUniformConvolver* cv[2] = { cl, cr };
        for (int c = 0; c < (R ? 2 : 1); ++c)
        {
            if (!cv[c]) continue;
            cv[c]->Add(ch[c], n);
            const int got = cv[c]->Get(ch[c], n);
            if (got < n) std::memset(ch[c] + std::max(got, 0), 0, sizeof(double) * (size_t) (n - std::max(got, 0)));
            if (outer)
                for (int i = 0; i < n; ++i)
                {
                    double v = ch[c][i];
                    if (!(std::isfinite(v) && std::fabs(v) < 1.0e300)) v = 0.0;
                    ch[c][i] = v * wetG;
                }
        }
        if (e)
        {
            juce::dsp::AudioBlock<double> blk(ch, R ? 2u : 1u, (size_t) n);
            e->proc.process(blk, e->params, e->cache.get());
        }
        if (epilogue)
            for (int c = 0; c < (R ? 2 : 1); ++c)
                for (int i = 0; i < n; ++i) ch[c][i] = (ch[c][i] * makeup) * 0.8912509381337456;
    }
}

// ---- output stages: the reference's own OutputFilter (src/OutputFilter.cpp) and UltraHighRateDCBlocker, driven per
// callback in the order of DSPCore::processDouble (DSPCoreDouble.cpp:452-469) and processOutputDouble (:600-602,
// 655-737, SimplePeakLimiter left out); the headroom multiply, scrub and clamp loops are three-liners restated here
// because AudioEngine.Processing.DSPCoreDouble.cpp itself needs the whole JUCE application to compile.
struct RefOut
{
    convo::OutputFilter filt;
    convo::UltraHighRateDCBlocker dcL, dcR;
    SimplePeakLimiter limiter;   // audioengine/SimplePeakLimiter.h, between the scrub and the hard clamp (DSPCoreDouble.cpp:700-710)
    bool limiterOn = false;
    double sr = 48000.0;
};
void* cpqref_out_create(double sr, double dc_cutoff)
{
    auto* o = new RefOut;
    o->filt.prepare(sr);
    o->dcL.init(sr, dc_cutoff);
    o->dcR.init(sr, dc_cutoff);
    o->sr = sr;
    return o;
}
void cpqref_out_set_limiter(void* h, double release_ms)
{
    auto* o = static_cast<RefOut*>(h);
    o->limiterOn = release_ms > 0.0;
    o->limiter.prepare(o->sr, release_ms);   // the engine: 100 ms (AudioEngine.Processing.DSPCoreLifecycle.cpp:228)
    o->limiter.reset();
}
void cpqref_out_destroy(void* h) { delete static_cast<RefOut*>(h); }
void cpqref_out_design(double sr, int conv_is_last, int hc, int lc, int lp, double* out15)
{
    convo::OutputFilter f;
    f.prepare(sr);
    const convo::BiquadCoeff* st[3];
    if (conv_is_last) { st[0] = &f.lcCoeff[lc]; st[1] = &f.hcCoeff[hc][0]; st[2] = &f.hcCoeff[hc][1]; }
    else { st[0] = &f.hpfCoeff; st[1] = &f.lpCoeff[lp][0]; st[2] = &f.lpCoeff[lp][1]; }
    for (int i = 0; i < 3; ++i)
    {
        out15[5 * i] = st[i]->b0; out15[5 * i + 1] = st[i]->b1; out15[5 * i + 2] = st[i]->b2;
        out15[5 * i + 3] = st[i]->a1; out15[5 * i + 4] = st[i]->a2;
    }
}
void cpqref_out_process(void* h, double* L, double* R, long total, int block, int use_filter, int conv_is_last, int hc, int lc,
                        int lp, double makeup, int use_dc, int headroom, int clamp)
{
    auto* o = static_cast<RefOut*>(h);
    juce::ScopedNoDenormals nd;
    constexpr double kOutputHeadroom = 0.8912509381337456;
    for (long pos = 0; pos < total; pos += block)
    {
        const int n = (int) std::min<long>(block, total - pos);
        double* ch[2] = { L + pos, R ? R + pos : nullptr };
        const int nch = R ? 2 : 1;
        if (use_filter)
        {
            juce::dsp::AudioBlock<double> blk(ch, (size_t) nch, (size_t) n);
            o->filt.process(blk, conv_is_last != 0, static_cast<convo::HCMode>(hc), static_cast<convo::LCMode>(lc), static_cast<convo::HCMode>(lp));
        }
        for (int c = 0; c < nch; ++c)
            for (int i = 0; i < n; ++i) ch[c][i] *= makeup;
        if (use_dc) o->dcL.processStereo(ch[0], ch[1], n, o->dcR);
        for (int c = 0; c < nch; ++c)
        {
            if (headroom)
                for (int i = 0; i < n; ++i) ch[c][i] *= kOutputHeadroom;
            if (clamp)
                for (int i = 0; i < n; ++i)
                {
                    double v = ch[c][i];
                    if (!(std::isfinite(v) && std::fabs(v) < 1.0e300)) v = 0.0;
                    ch[c][i] = v;
                }
        }
        if (o->limiterOn) o->limiter.processBlock(ch[0], ch[1], n, 0.8413951287507587, 0.108748);   // kPLThreshold, kPLKnee
        if (clamp)
            for (int c = 0; c < nch; ++c)
                for (int i = 0; i < n; ++i) ch[c][i] = std::min(std::max(ch[c][i], -kOutputHeadroom), kOutputHeadroom);
    }
}

// IRAnalyzer::estimateMaxFrequencyResponseGain (src/IRAnalyzer.cpp:63-155), the FFT stage of IRConverter::computeScaleFactor
// (IRConverter.cpp itself needs JUCE's audio-format classes and is not compiled).
double cpqref_ir_freq_peak_gain(const double* l, const double* r, int n)
{
    double* ch[2] = { const_cast<double*>(l), const_cast<double*>(r) };
    juce::AudioBuffer<double> buf(ch, r ? 2 : 1, n);
    return IRAnalyzer::estimateMaxFrequencyResponseGain(buf);
}

// convo::input_transform::convertDoubleToDoubleHighQuality (src/InputBitDepthTransform.h:123-133): the engine's input stage
// (DSPCore::processInput, AudioEngine.Processing.DSPCoreIO.cpp:203-232) on a double buffer, in place.
void cpqref_input_transform(double* data, int n, double gain)
{
    convo::input_transform::convertDoubleToDoubleHighQuality(data, data, n, gain);
}

// The same with the VSL generator unavailable (vslNewStream fails): every uniform comes from the header's own
// fallbackUniform (xorshift64*, :485-497), seeded per channel through SplitMix64(seed) (:118-140) -- a mode of the reference
// that needs no injected numbers, so the library's cpq_set_dither_seed is pinned against it bit for bit.
void cpqref_dither_process_fallback(double* L, double* R, long total, int block, double sr, int bits, double headroom,
                                    unsigned long long seed, double* z_out)
{
    cpqref_vsl_begin();
    g_vslFail = true;
    auto* d = new convo::PsychoacousticDither(std::optional<uint64_t>(seed));
    d->prepare(sr, bits);
    for (long pos = 0; pos < total; pos += block)
    {
        const int n = (int) std::min<long>(block, total - pos);
        d->processStereoBlock(L + pos, R ? R + pos : nullptr, n, headroom);
        d->refillRandomRingNonRt();
    }
    if (z_out)
        for (int c = 0; c < 2; ++c)
            for (int t = 0; t < 12; ++t) z_out[c * 12 + t] = d->shaperStateBuffer[c * convo::PsychoacousticDither::STATE_STRIDE + t];
    delete d;
    g_vslFail = false;
    cpqref_vsl_begin();
}

// UltraHighRateDCBlocker::init + process on one buffer (the IR loader runs it at 1 Hz, LoaderThread.cpp:590-598)
void cpqref_ir_dc_block(double* data, int n, double sr, double cutoff)
{
    convo::UltraHighRateDCBlocker dc;
    dc.init(sr, cutoff);
    dc.process(data, n);
}

int cpqref_abi_version(void) { return 5; }

// ------------------------------------------------------------------ dither (PsychoacousticDither.h, unmodified)
// The header's only MKL dependency is the VSL uniform generator (:79,418-431); ref_shim/mkl_vsl.h replaces it with
// streams that hand out injected numbers, so the noise shaper / quantiser recurrence of processStereoBlock (:293-405)
// runs exactly as compiled from the reference.  uL / uR: 2 * total uniforms per channel (u1, u2 per sample), consumed in
// the order the ring delivers them (prefill of 65536 at prepare(), refillRandomRingNonRt between callbacks, :440-475).
// z_out (nullable): shaperStateBuffer of channels 0 and 1 after the run, [2][12].
void cpqref_dither_process(double* L, double* R, long total, int block, double sr, int bits, double headroom,
                           const double* uL, const double* uR, double* z_out)
{
    cpqref_vsl_begin();
    cpqref_vsl_inject(0, uL, 2 * total);
    if (R) cpqref_vsl_inject(1, uR, 2 * total);
    auto* d = new convo::PsychoacousticDither(std::optional<uint64_t>(1));
    d->prepare(sr, bits);
    for (long pos = 0; pos < total; pos += block)
    {
        const int n = (int) std::min<long>(block, total - pos);
        d->processStereoBlock(L + pos, R ? R + pos : nullptr, n, headroom);
        // the worker thread's job (AudioEngine's refill timer): keep the ring topped up so it never runs dry
        for (int k = 0; k < 2 * n / 2048 + 2; ++k) d->refillRandomRingNonRt();
    }
    if (z_out)
        for (int c = 0; c < 2; ++c)
            for (int t = 0; t < 12; ++t) z_out[c * 12 + t] = d->shaperStateBuffer[c * convo::PsychoacousticDither::STATE_STRIDE + t];
    delete d;
    cpqref_vsl_begin();
}
}
