// cpq.hpp -- header-only C++20 host wrapper over the C ABI (include/cpq.h).
//
// Method names follow the reference classes this path replaces, in their offline / batched form:
//
//   convo::MKLNonUniformConvolver::SetImpulse / Reset / getLatency      (src/MKLNonUniformConvolver.h:197-242)
//   StereoConvolver::init(irL, irR, length, sr, ..., scale, ..., FilterSpec*)   (src/ConvolverProcessor.h:741-744)
//   ConvolverProcessor::prepareToPlay(sr, samplesPerBlock)              (src/convolver/ConvolverProcessor.Lifecycle.cpp:211)
//   EQProcessor::prepareToPlay / setBandFrequency / setBandGain / setBandQ / setBandType / setBandEnabled /
//                setBandChannelMode / setTotalGain / setNonlinearSaturation / process
//                                                                        (src/eqprocessor/EQProcessor.h:196-260)
//   DSPCore::processDouble order ConvolverThenEQ + outputMakeupGain + kOutputHeadroom / dither
//                                                                        (src/audioengine/AudioEngine.Processing.DSPCoreDouble.cpp:386-469,577-663)
//
// Error behaviour: like the reference's noexcept audio path, nothing throws from process(); every call returns
// bool / cpq_status and lastError() gives the reason.  There is no CPU fallback.
#pragma once

#include <array>
#include <cstdint>
#include <algorithm>
#include <span>
#include <string>
#include <utility>
#include <vector>

#include "cpq.h"

namespace convopeq_b200
{

// convo::EQBandParams (src/core/EQParameters.h:11-20): float parameters, promoted to double by the designer
struct EQBandParams
{
    float frequency = 1000.0f;
    float gain = 0.0f;
    float q = 0.707f;
    bool enabled = true;
    int type = 1;         // 0 LowShelf, 1 Peaking, 2 HighShelf, 3 LowPass, 4 HighPass
    int channelMode = 0;  // 0 Stereo, 1 Left, 2 Right
};

// convo::EQParameters (src/core/EQParameters.h:23-49); band channelMode 0 Stereo, 1 Left, 2 Right, 3 Mid, 4 Side
struct EQParameters
{
    std::array<EQBandParams, CPQ_NUM_BANDS> bands {};
    float totalGainDb = 0.0f;
    bool agcEnabled = false;
    float nonlinearSaturation = 0.2f;
    int filterStructure = 0;   // 0 Serial, 1 Parallel

    EQParameters()
    {
        constexpr float f[CPQ_NUM_BANDS] = { 20, 32, 50, 80, 125, 200, 315, 500, 800, 1250, 2000, 3150, 5000, 8000, 12500, 16000, 19000, 20000, 22000, 24000 };
        for (int i = 0; i < CPQ_NUM_BANDS; ++i) bands[(size_t) i].frequency = f[i];
    }
};

using FilterSpec = cpq_filter_spec;   // convo::FilterSpec; defaults through defaultFilterSpec()

inline FilterSpec defaultFilterSpec()
{
    FilterSpec s {};
    cpq_filter_spec_default(&s);
    return s;
}

/// A batch of independent stereo (or mono) streams on one B200: convolver -> EQ -> output epilogue.
class BatchEngine
{
public:
    BatchEngine() = default;
    BatchEngine(const BatchEngine&) = delete;
    BatchEngine& operator=(const BatchEngine&) = delete;
    BatchEngine(BatchEngine&& o) noexcept : h_(std::exchange(o.h_, nullptr)), cfg_(o.cfg_), eq_(std::move(o.eq_)) {}
    ~BatchEngine() { releaseResources(); }

    /// ConvolverProcessor::prepareToPlay + EQProcessor::prepareToPlay for every stream of the batch.
    bool prepareToPlay(double sampleRate, int samplesPerBlock, int numStreams, int numChannels, std::int64_t maxSamples,
                       int device = 0, bool convolverOuterBoundary = true, bool sharedIR = false, bool sharedEQ = false)
    {
        releaseResources();
        cpq_config_default(&cfg_);
        cfg_.device = device;
        cfg_.n_streams = numStreams;
        cfg_.n_channels = numChannels;
        cfg_.block_size = samplesPerBlock;
        cfg_.sample_rate = sampleRate;
        cfg_.max_samples = maxSamples;
        cfg_.conv_boundary = convolverOuterBoundary ? CPQ_CONV_OUTER : CPQ_CONV_INNER;
        cfg_.shared_ir = sharedIR ? 1 : 0;
        cfg_.shared_eq = sharedEQ ? 1 : 0;
        status_ = cpq_create(&cfg_, &h_);
        if (status_ != CPQ_OK) createError_ = cpq_last_error(nullptr);
        eq_.assign((size_t) (sharedEQ ? 1 : numStreams), EQParameters {});
        return status_ == CPQ_OK;
    }

    void releaseResources() noexcept
    {
        if (h_) cpq_destroy(h_);
        h_ = nullptr;
    }

    /// MKLNonUniformConvolver::SetImpulse for one stream-channel (stream = -1 with a shared IR). `impulse` is borrowed.
    bool SetImpulse(int stream, int channel, std::span<const double> impulse, double scale = 1.0, const FilterSpec* filterSpec = nullptr)
    {
        return ok(cpq_set_impulse(h_, stream, channel, impulse.data(), (int) impulse.size(), scale, filterSpec));
    }

    /// StereoConvolver::init: both channels at once (does not take ownership, unlike the reference).
    bool init(int stream, std::span<const double> irL, std::span<const double> irR, double scale = 1.0, const FilterSpec* filterSpec = nullptr)
    {
        return SetImpulse(stream, 0, irL, scale, filterSpec) && (cfg_.n_channels < 2 || SetImpulse(stream, 1, irR, scale, filterSpec));
    }

    /// ConvolverProcessor::loadImpulseResponse (src/ConvolverProcessor.h:235) for one stream, from the bytes of a WAV file at the
    /// engine's sample rate: the loader thread's decode / trim / window / phase mode / scale factor / peak latency steps
    /// (convolver/ConvolverProcessor.LoaderThread.cpp:430-757), then init.  phaseMode: 0 AsIs, 1 Minimum, 2 Mixed.
    bool loadImpulseResponse(int stream, std::span<const std::uint8_t> wavFile, int phaseMode = 0, double targetIRLengthSec = 1.0,
                             const FilterSpec* filterSpec = nullptr, cpq_ir_load_info* info = nullptr)
    {
        return ok(cpq_load_impulse_wav(h_, stream, wavFile.data(), wavFile.size(), phaseMode, targetIRLengthSec, filterSpec, info));
    }

    // ---- EQProcessor parameter setters (values are float like the reference's UI parameters) ----
    void setBandFrequency(int stream, int band, float f) { if (valid(stream, band)) { p(stream).bands[(size_t) band].frequency = f; dirty_ = true; } }
    void setBandGain(int stream, int band, float g) { if (valid(stream, band)) { p(stream).bands[(size_t) band].gain = g; dirty_ = true; } }
    void setBandQ(int stream, int band, float q) { if (valid(stream, band)) { p(stream).bands[(size_t) band].q = q; dirty_ = true; } }
    void setBandType(int stream, int band, int type) { if (valid(stream, band)) { p(stream).bands[(size_t) band].type = type; dirty_ = true; } }
    void setBandEnabled(int stream, int band, bool e) { if (valid(stream, band)) { p(stream).bands[(size_t) band].enabled = e; dirty_ = true; } }
    void setBandChannelMode(int stream, int band, int m) { if (valid(stream, band)) { p(stream).bands[(size_t) band].channelMode = m; dirty_ = true; } }
    /// EQProcessor::setTotalGain: jlimit(DSP_MIN_GAIN_DB, DSP_MAX_GAIN_DB) = +-48 dB (EQProcessor.Parameters.cpp:103-106)
    void setTotalGain(int stream, float db) { if (valid(stream, 0)) { p(stream).totalGainDb = std::min(48.0f, std::max(-48.0f, db)); dirty_ = true; } }
    /// EQProcessor::setNonlinearSaturation: jlimit(0, 1) (EQProcessor.Parameters.cpp:206-208)
    void setNonlinearSaturation(int stream, float s) { if (valid(stream, 0)) { p(stream).nonlinearSaturation = std::min(1.0f, std::max(0.0f, s)); dirty_ = true; } }
    /// EQProcessor::loadFromTextFile on the contents of an EqualizerAPO / AutoEq preset (EQProcessor.Core.cpp:300-495).
    bool loadFromText(int stream, const std::string& text)
    {
        if (!valid(stream, 0)) return false;
        cpq_eq_band_params b[CPQ_NUM_BANDS];
        EQParameters& q = p(stream);
        for (int i = 0; i < CPQ_NUM_BANDS; ++i)
            b[i] = { q.bands[(size_t) i].frequency, q.bands[(size_t) i].gain, q.bands[(size_t) i].q, q.bands[(size_t) i].enabled ? 1 : 0,
                     q.bands[(size_t) i].type, q.bands[(size_t) i].channelMode };
        if (cpq_parse_eq_preset(text.c_str(), b, &q.totalGainDb) < 0) return false;
        for (int i = 0; i < CPQ_NUM_BANDS; ++i)
            q.bands[(size_t) i] = { b[i].frequency, b[i].gain_db, b[i].q, b[i].enabled != 0, b[i].type, b[i].channel_mode };
        dirty_ = true;
        return true;
    }
    void setEQParameters(int stream, const EQParameters& params) { if (valid(stream, 0)) { p(stream) = params; dirty_ = true; } }

    /// outputMakeupGain + dither bit depth (0 = the no-dither kOutputHeadroom branch).
    bool setOutputStage(double makeupGain, int ditherBitDepth = 0) { return ok(cpq_set_epilogue(h_, makeupGain, ditherBitDepth)); }
    /// convo::OutputFilter::process(block, convIsLast, hcMode, lcMode, lpMode) (OutputFilter.h:108-131): runs when
    /// CPQ_STAGE_OUTPUT_FILTER is requested.  Mode values as convo::HCMode / convo::LCMode.
    bool setOutputFilter(bool enabled, bool convIsLast, int hcMode = 1, int lcMode = 0, int lpMode = 1)
    {
        return ok(cpq_set_output_filter(h_, enabled ? 1 : 0, convIsLast ? 1 : 0, hcMode, lcMode, lpMode));
    }
    /// Output DC blocker (the engine initialises it at 3 Hz, AudioEngine.h:643-651; 0 = off) and processOutputDouble's
    /// scrub + +-kOutputHeadroom hard clamp, both inside CPQ_STAGE_EPILOGUE.
    /// convolverInputTrimGain (EQThenConvolver order: pass CPQ_ORDER_EQ_THEN_CONV in `stages`).
    bool setConvolverInputTrim(double gain) { return ok(cpq_set_conv_input_trim(h_, gain)); }
    /// Gain of DSPCore::processInput's input transform (CPQ_STAGE_INPUT).
    bool setInputGain(double gain) { return ok(cpq_set_input_gain(h_, gain)); }
    /// SimplePeakLimiter (release 100 ms in the reference engine) between the scrub and the hard clamp; 0 = off.
    bool setPeakLimiter(double releaseMs) { return ok(cpq_set_peak_limiter(h_, releaseMs)); }
    /// ConvolverProcessor::setBypass: the convolver stage only delays by the dry path's latency compensation.
    bool setConvolverBypass(bool bypassed) { return ok(cpq_set_convolver_bypass(h_, bypassed ? 1 : 0)); }
    /// enableDirectHead of SetImpulse / StereoConvolver::init; call before the first SetImpulse.
    bool setDirectHeadEnabled(bool enable) { return ok(cpq_set_direct_head(h_, enable ? 1 : 0)); }
    /// ConvolverProcessor::setMix (Runtime.cpp:816) + the dry path's latency compensation (settled state).
    bool setMix(float mix, int dryDelaySamples) { return ok(cpq_set_mix(h_, mix, dryDelaySamples)); }
    bool setOutputProtection(double dcCutoffHz, bool hardClamp) { return ok(cpq_set_output_stage(h_, dcCutoffHz, hardClamp ? 1 : 0)); }
    /// PsychoacousticDither(seed) per stream, uniforms from the header's own fallback generator (no injected numbers needed).
    bool setDitherSeeds(std::span<const std::uint64_t> streamSeeds)
    {
        return (int) streamSeeds.size() == cfg_.n_streams && ok(cpq_set_dither_seed(h_, streamSeeds.data()));
    }
    bool setDitherUniforms(std::span<const double> uniforms, std::int64_t samplesPerChannel)
    {
        return ok(cpq_set_dither_uniforms(h_, uniforms.data(), samplesPerChannel));
    }

    /// In place on planar host buffers, planar[stream * numChannels + ch] -> numSamples doubles.
    bool process(double* const* planar, std::int64_t numSamples, unsigned stages = CPQ_STAGE_ALL)
    {
        if ((stages & CPQ_STAGE_EQ) && !pushEq()) return false;
        return ok(cpq_process(h_, planar, numSamples, stages));
    }

    /// The same with float host buffers (FP32 on the wire, FP64 arithmetic).
    bool process(float* const* planar, std::int64_t numSamples, unsigned stages = CPQ_STAGE_ALL)
    {
        if ((stages & CPQ_STAGE_EQ) && !pushEq()) return false;
        return ok(cpq_process_f32(h_, planar, numSamples, stages));
    }

    /// In place on device memory [numStreams * numChannels][stride].
    bool processDevice(double* deviceIO, std::int64_t stride, std::int64_t numSamples, unsigned stages = CPQ_STAGE_ALL)
    {
        if ((stages & CPQ_STAGE_EQ) && !pushEq()) return false;
        return ok(cpq_process_device(h_, deviceIO, stride, numSamples, stages));
    }

    bool Reset() { return ok(cpq_reset(h_)); }
    /// Streaming continuation: process() calls continue the stream (Add/Get/EQ state carried) until Reset().
    bool setStreaming(bool enable) { return ok(cpq_set_streaming(h_, enable ? 1 : 0)); }
    std::int64_t streamPosition() const { return cpq_stream_position(h_); }
    std::vector<unsigned char> exportState()
    {
        std::vector<unsigned char> blob(cpq_state_size(h_));
        if (blob.empty() || !ok(cpq_export_state(h_, blob.data(), blob.size()))) blob.clear();
        return blob;
    }
    bool importState(const std::vector<unsigned char>& blob) { return ok(cpq_import_state(h_, blob.data(), blob.size())); }
    int getLatency() const { return cpq_latency(h_); }
    bool isReady() const { return h_ != nullptr; }
    cpq_layout getLayout() const { cpq_layout l {}; if (h_) cpq_get_layout(h_, &l); return l; }
    cpq_timings getTimings() const { cpq_timings t {}; if (h_) cpq_get_timings(h_, &t); return t; }
    cpq_status lastStatus() const { return status_; }
    std::string lastError() const { return h_ ? std::string(cpq_last_error(h_)) : createError_; }
    cpq_handle handle() const { return h_; }

private:
    bool ok(cpq_status s) { status_ = s; return s == CPQ_OK; }
    bool valid(int stream, int band) const
    {
        const int n = (int) eq_.size();
        return h_ && band >= 0 && band < CPQ_NUM_BANDS && ((cfg_.shared_eq && (stream == -1 || stream == 0)) || (stream >= 0 && stream < n));
    }
    EQParameters& p(int stream) { return eq_[cfg_.shared_eq ? 0 : (size_t) stream]; }

    // EQProcessor::createCoeffCache (ProcessingCache.cpp:56-96): design every enabled band, hand the cache to the engine
    bool pushEq()
    {
        if (!dirty_) return true;
        for (size_t s = 0; s < eq_.size(); ++s)
        {
            cpq_svf_coeffs co[CPQ_NUM_BANDS] {};
            std::uint8_t active[CPQ_NUM_BANDS] {};
            std::int32_t mode[CPQ_NUM_BANDS] {};
            std::uint8_t nodeActive[CPQ_NUM_BANDS] {};
            for (int b = 0; b < CPQ_NUM_BANDS; ++b)
            {
                const EQBandParams& bp = eq_[s].bands[(size_t) b];
                active[b] = (bp.enabled && cfg_.sample_rate > 0.0) ? 1 : 0;
                nodeActive[b] = (std::uint8_t) cpq_band_node_active(bp.type, bp.gain, bp.enabled ? 1 : 0, cfg_.sample_rate);
                mode[b] = bp.channelMode;
                if (active[b] && cpq_design_band(bp.type, bp.frequency, bp.gain, bp.q, cfg_.sample_rate, &co[b]) != CPQ_OK) return ok(CPQ_ERR_INVALID);
            }
            const double sat = static_cast<double>(eq_[s].nonlinearSaturation);   // float -> double like Processing.cpp:1114
            const int st = cfg_.shared_eq ? -1 : (int) s;
            if (!ok(cpq_set_eq(h_, st, co, active, mode, sat, cpq_db_to_gain(eq_[s].totalGainDb)))) return false;
            if (!ok(cpq_set_eq_mode(h_, st, eq_[s].filterStructure, eq_[s].agcEnabled ? 1 : 0, nodeActive))) return false;
        }
        dirty_ = false;
        return true;
    }

    cpq_handle h_ = nullptr;
    cpq_config cfg_ {};
    std::vector<EQParameters> eq_;
    bool dirty_ = true;
    cpq_status status_ = CPQ_OK;
    std::string createError_;
};

} // namespace convopeq_b200
