/*
 * cpq.h -- C ABI of convopeq_b200: the B200 (sm_100a) implementation of ConvoPeq's DSP hot path
 * in offline / batched form:
 *
 *   FP64 non-uniform partitioned overlap-save convolver  ->  20-band TPT-SVF EQ  ->  gain / dither
 *
 * This header is the drop-in boundary.  Each entry point names the reference interface it
 * replaces (paths relative to the reference's src/).  Plain C types only: no torch, no C++.
 * One handle = one CUDA device + one batch of independent stereo (or mono) streams that share a
 * sample rate, a host block size and a layer geometry.  A handle is thread-compatible, not thread-safe: one
 * thread at a time may call into a given handle (any thread; every entry point selects the handle's device itself);
 * different handles, on the same or on different devices, are independent and may be driven concurrently.  Errors are status codes (the reference's path is noexcept and reports
 * failure as silence + counters); cpq_last_error() gives the text.  There is no CPU fallback:
 * every compute entry point fails with CPQ_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef CPQ_H
#define CPQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPQ_ABI_VERSION 5
#define CPQ_NUM_BANDS 20        /* EQProcessor::NUM_BANDS, eqprocessor/EQProcessor.h:153 */
#define CPQ_MAX_LAYERS 3        /* MKLNonUniformConvolver::kNumLayers, MKLNonUniformConvolver.h:391 */
#define CPQ_NS_ORDER 12         /* PsychoacousticDither::NS_ORDER, PsychoacousticDither.h:60 */

typedef struct cpq_engine* cpq_handle;

typedef enum cpq_status
{
    CPQ_OK = 0,
    CPQ_ERR_INVALID = 1,      /* bad argument (the reference returns false / silently ignores) */
    CPQ_ERR_NOT_READY = 2,    /* process before every stream-channel has an impulse / EQ */
    CPQ_ERR_CUDA = 3,         /* CUDA runtime error or no usable device */
    CPQ_ERR_OOM = 4,
    CPQ_ERR_UNSUPPORTED = 5,  /* reference feature outside the hot path (host blocks that are not a power of two ...) */
    CPQ_ERR_GEOMETRY = 6      /* stream-channels of one handle must share the layer geometry */
} cpq_status;

/* convo::FilterSpec, MKLNonUniformConvolver.h:123-133 (same fields, same defaults when zero-initialised
 * through cpq_filter_spec_default). */
typedef struct cpq_filter_spec
{
    double sample_rate;          /* 48000 */
    int32_t hc_mode;             /* convo::HCMode  0 Sharp, 1 Natural (default), 2 Soft   OutputFilter.h:75 */
    int32_t lc_mode;             /* convo::LCMode  0 Natural (default), 1 Soft             OutputFilter.h:85 */
    int32_t tail_mode;           /* 0 air absorption, 1 layer tail contouring (default), 2 bypass */
    int32_t tail_enabled;        /* 1 */
    double tail_start_seconds;   /* 0.085 */
    double tail_strength;        /* 1.0 */
    int32_t tail_l1l2_multiplier;/* 8 */
    int32_t reserved_;
} cpq_filter_spec;

/* EQCoeffsSVF a1,a2,a3,m0,m1,m2 (eqprocessor/EQProcessor.h:91-96) */
typedef struct cpq_svf_coeffs
{
    double a1, a2, a3, m0, m1, m2;
} cpq_svf_coeffs;

/* Which reference boundary cpq_process reproduces for the convolver stage. */
typedef enum cpq_conv_boundary
{
    CPQ_CONV_INNER = 0, /* MKLNonUniformConvolver Add/Get (StereoConvolver::process, ConvolverProcessor.Runtime.cpp:1159) */
    CPQ_CONV_OUTER = 1  /* ConvolverProcessor::process at mix = 1: scrub(|x|>=1e300 or non-finite -> 0) then
                           x * equalPowerSin(1.0) (Runtime.cpp:26-31,50-60,675,722,748) */
} cpq_conv_boundary;

/* cpq_process stage selection (DSPCore::processDouble order ConvolverThenEQ, AudioEngine.Processing.DSPCoreDouble.cpp:386-414) */
enum
{
    CPQ_STAGE_CONV = 1u,      /* convolverRt().process */
    CPQ_STAGE_EQ = 2u,        /* eqRt().process(block, params, cache) */
    CPQ_STAGE_EPILOGUE = 4u,  /* makeup gain (:465-469) + processOutputDouble headroom / dither (:581,:644-663) */
    CPQ_STAGE_ALL = 7u,       /* the north-star path: conv -> EQ -> gain / dither */
    CPQ_STAGE_OUTPUT_FILTER = 8u, /* outputFilter.process between the EQ and the makeup gain (:460-462), see cpq_set_output_filter */
    CPQ_STAGE_FULL = 15u,
    /* ProcessingOrder::EQThenConvolver (:415-451) instead of ConvolverThenEQ: EQ -> convolverInputTrimGain -> convolver
     * -> (output filter with conv_is_last) -> epilogue.  Only meaningful together with CPQ_STAGE_CONV | CPQ_STAGE_EQ. */
    CPQ_ORDER_EQ_THEN_CONV = 16u,
    /* DSPCore::processInput's transform of the input block before everything else (convo::input_transform::
     * applyHighQuality64BitTransform, InputBitDepthTransform.h:86-100; AudioEngine.Processing.DSPCoreIO.cpp:203-232): input gain
     * (cpq_set_input_gain, the reference passes 1.0), NaN and |v| < 1e-20 -> 0, clamp to [-1, 1].  Not part of CPQ_STAGE_FULL. */
    CPQ_STAGE_INPUT = 32u
};

typedef struct cpq_config
{
    int32_t device;        /* CUDA ordinal */
    int32_t n_streams;     /* independent streams in the batch */
    int32_t n_channels;    /* channels per stream: 1 or 2 (the reference engine is <= 2 channels) */
    int32_t block_size;    /* host callback size the reference would run with: 64..8192.  A power of two is the regular
                              case; otherwise (480, 441 ...) the convolver is prepared with the block rounded up
                              to a power of two and called with this block, like the application does
                              (ConvolverProcessor.LoaderThread.cpp:230,239-245) */
    double sample_rate;
    int64_t max_samples;   /* capacity per channel for one cpq_process call (multiple of block_size) */
    int32_t conv_boundary; /* cpq_conv_boundary */
    int32_t shared_ir;     /* 1: one IR pair (per channel) shared by all streams */
    int32_t shared_eq;     /* 1: one EQ setting shared by all streams */
    int32_t uniform_partitions; /* 1: EXTENSION (BASELINE config 1 "uniform partitioned convolution"): the whole IR as one
                                   layer of ceil(len / block) uniform partitions, which the reference cannot express beyond
                                   32 partitions (kL0MaxParts); checked against linear convolution, not the reference */
    size_t workspace_bytes;/* upper bound for the per-call spectra workspace; 0 = default (free memory / 4 within 4..16 GiB) */
} cpq_config;

/* cpq_get_layout: what SetImpulse decided (MKLNonUniformConvolver.cpp:738-758,988-994,1004-1024) plus the
 * effective tail delay D = first output sample at which the layer contributes. */
typedef struct cpq_layer_layout
{
    int32_t part_size, fft_size, num_parts_ir, num_parts, parts_per_callback, output_delay_samples;
    int32_t ir_offset, ir_len;
    int64_t first_output_sample;  /* -1 if the layer never contributes within max_samples */
    int64_t skipped_callbacks;    /* callbacks in which delayLineReadAdd skipped this layer after it first spoke */
    double gain;                  /* m_tailLayerGain */
} cpq_layer_layout;

typedef struct cpq_layout
{
    int32_t num_layers;
    int32_t latency;              /* getLatency(): L0 partition size (reported, not the real onset) */
    cpq_layer_layout layers[CPQ_MAX_LAYERS];
} cpq_layout;

typedef struct cpq_timings
{
    float h2d_ms, fft_fwd_ms, mac_ms, fft_inv_ms, eq_ms, d2h_ms, total_ms;
    int32_t kernel_launches;
    int32_t chunks;           /* sequence chunks of the last call = launches of each stage kernel per layer */
} cpq_timings;

/* ---- lifecycle --------------------------------------------------------------------------- */
int cpq_abi_version(void);
const char* cpq_status_string(cpq_status s);
const char* cpq_last_error(cpq_handle h);
void cpq_filter_spec_default(cpq_filter_spec* out);   /* FilterSpec{} defaults */
void cpq_config_default(cpq_config* out);

/* ConvolverProcessor::prepareToPlay(sr, block) + EQProcessor::prepareToPlay(sr, maxBlock)
 * (convolver/ConvolverProcessor.Lifecycle.cpp:211, eqprocessor/EQProcessor.Core.cpp:679). */
cpq_status cpq_create(const cpq_config* cfg, cpq_handle* out);
void cpq_destroy(cpq_handle h);

/* MKLNonUniformConvolver::Reset (MKLNonUniformConvolver.cpp:1693) + EQ / output-stage state clear + dither state clear.
 * In the default one-shot mode every process call starts from this state anyway; in streaming mode (below) this is the
 * call that starts a new stream. */
cpq_status cpq_reset(cpq_handle h);

/* Streaming continuation: the reference's path is a stream -- Add/Get keep the FDL, the partial input frame, the
 * accumulators and the delay-line cursors of every layer between callbacks (MKLNonUniformConvolver.h:288-365, .cpp:1407-1548;
 * cleared only by Reset, :1693), EQProcessor keeps filterState (EQProcessor.h:637) and the AGC envelopes, the limiter its
 * envelope, the dither its error history.  With enable = 1 every cpq_process / cpq_process_f32 / cpq_process_device call
 * continues where the previous one stopped: a signal processed in segments of any whole number of callbacks gives the
 * convolver output of the one-shot call bit for bit, and the EQ / output-stage output to rounding (the blocked scan's tiles
 * start at the segment boundary; identical bits when the segments are multiples of 8192 samples).  Carried per
 * stream-channel: the last 2 Pmax input samples, Q - 1 input spectra per layer, the tail samples Get has not read yet, and
 * the stage states.  cpq_reset returns to the Reset state; enable = 0 (default) makes every call start from Reset again.
 * Host blocks that are not a power of two (480, 441 ...) are carried through the reference's layer-0 output ring.
 * Not covered (CPQ_ERR_UNSUPPORTED from the process call): partition
 * ranges / stream windows, plans that drop tail blocks.  cpq_schedule_total_gain counts at_callback from the start of the call
 * that follows it; a ramp still in progress at the end of a call goes on in the next one (it is not part of the exported state:
 * export once it has settled). */
cpq_status cpq_set_streaming(cpq_handle h, int enable);
int64_t cpq_stream_position(cpq_handle h);            /* samples per channel processed since Reset (streaming mode) */
/* The carried state as one host blob (chaining segments across handles / processes / GPUs): export after any
 * call, import into a handle with the same configuration, impulses and EQ settings (CPQ_ERR_GEOMETRY otherwise), continue. */
size_t cpq_state_size(cpq_handle h);                  /* 0 when the handle is not in streaming mode */
cpq_status cpq_export_state(cpq_handle h, void* dst, size_t bytes);
cpq_status cpq_import_state(cpq_handle h, const void* src, size_t bytes);

/* ---- prepare ----------------------------------------------------------------------------- */
/* MKLNonUniformConvolver::SetImpulse(impulse, irLen, blockSize, scale, enableDirectHead=false, filterSpec)
 * (MKLNonUniformConvolver.h:197-200, .cpp:610-1149).  `ir` is borrowed (copied).  spec == NULL is the
 * reference's nullptr.  stream = -1 with cfg.shared_ir sets the channel's IR for every stream. */
cpq_status cpq_set_impulse(cpq_handle h, int stream, int channel, const double* ir, int ir_len, double scale,
                           const cpq_filter_spec* spec);

/* enableDirectHead of SetImpulse / StereoConvolver::init (MKLNonUniformConvolver.h:197-200, ConvolverProcessor.h:741-744;
 * "experimental" in the reference, default off): the first min(irLen, 32) taps run as a direct-form FIR
 * (processDirectBlock, MKLNonUniformConvolver.cpp:1169-1232) added to the L0 output in Get (:1604-1616), and are removed from
 * the impulse the partitions are built from (:730-731), so they also bypass the spectrum filter.  Applies to every
 * cpq_set_impulse that follows: call it before the first one. */
cpq_status cpq_set_direct_head(cpq_handle h, int enable);

/* EQCoeffCache + EQParameters as consumed by EQProcessor::process(block, params, cache)
 * (eqprocessor/EQProcessor.h:121-138, ProcessingCache.cpp:56-96, Processing.cpp:1019-1276); Serial structure and AGC
 * off unless cpq_set_eq_mode says otherwise.  chan_mode: 0 Stereo, 1 Left, 2 Right, 3 Mid, 4 Side (EQChannelMode,
 * EQProcessor.h:55-62; an active Mid/Side band sends the reference to its node path, Processing.cpp:1037-1044, 690-740,
 * which this library reproduces on stereo handles, in the Serial and in the Parallel structure).
 * saturation is the already-promoted double, e.g. (double)0.2f.  total_gain_lin is
 * Decibels::decibelsToGain((double)totalGainDb) (settled LinearRamp). stream = -1 with cfg.shared_eq. */
cpq_status cpq_set_eq(cpq_handle h, int stream, const cpq_svf_coeffs coeffs[CPQ_NUM_BANDS],
                      const uint8_t active[CPQ_NUM_BANDS], const int32_t chan_mode[CPQ_NUM_BANDS],
                      double saturation, double total_gain_lin);

/* EQParameters::filterStructure (0 Serial, 1 Parallel: Processing.cpp:1132-1228, out = src + sum_b (band_b(src) - src))
 * and EQParameters::agcEnabled (block-rate AGC, Processing.cpp:1119-1131 + processAGC :343-445; replaces the total-gain
 * ramp).  node_active (nullable): BandNode::active per band as createBandNode computes it (EQProcessor.Coefficients.cpp:
 * 27-58, see cpq_band_node_active) -- the band set of the node path, consulted only when an active band is Mid/Side;
 * NULL = same as cpq_set_eq's `active`.  Call after cpq_set_eq (which does not change these).  stream = -1 with shared_eq. */
cpq_status cpq_set_eq_mode(cpq_handle h, int stream, int filter_structure, int agc_enabled, const uint8_t node_active[CPQ_NUM_BANDS]);
/* Host-only: createBandNode's activity rule: enabled, sample rate > 0, and not a shelf / peaking band with |gain| < 0.01 dB. */
int cpq_band_node_active(int type, float gain_db, int enabled, double sample_rate);
/* rtAgcEnvInputShadow, rtAgcEnvOutputShadow, rtAgcCurrentGainShadow after the last cpq_process (AGC streams). */
cpq_status cpq_get_agc_state(cpq_handle h, int stream, double out[3]);

/* EQProcessor::setTotalGain at a callback boundary (EQProcessor.Parameters.cpp:109): from callback index
 * `at_callback` of the next cpq_process the total gain ramps to new_gain_lin over 50 ms
 * (LinearRamp, DspNumericPolicy.h:319-421; Processing.cpp:1262-1274). */
cpq_status cpq_schedule_total_gain(cpq_handle h, int stream, int64_t at_callback, double new_gain_lin);

/* outputMakeupGain (DSPCoreDouble.cpp:465-469), kOutputHeadroom / dither (:581,:644-663,
 * PsychoacousticDither.h:192-355).  dither_bits <= 0: y *= makeup * 0.8912509381337456.
 * dither_bits > 0: uniforms must hold 2*T doubles per stream-channel, planar
 * [stream*n_channels + ch][2*T] (u1,u2 per sample) on the host; they are the injected replacement of the
 * reference's MKL VSL ring. */
cpq_status cpq_set_epilogue(cpq_handle h, double makeup_gain, int dither_bits);
cpq_status cpq_set_dither_uniforms(cpq_handle h, const double* uniforms, int64_t samples_per_channel);
/* Uniforms from the reference's own generator instead of injected ones: PsychoacousticDither falls back to a per-channel
 * xorshift64* generator whenever its VSL stream is not valid (fallbackUniform, PsychoacousticDither.h:485-497, used by
 * nextTPDF_MKL :560-572), seeded per channel through SplitMix64(seed) (:118-140).  stream_seeds[n_streams]: the `seed` argument
 * of each stream's PsychoacousticDither(seed); every process call then draws two uniforms per channel-sample on the device --
 * no 16 bytes per sample of host-supplied numbers -- and equals the reference bit for bit in that mode (pinned with a VSL
 * shim whose vslNewStream fails).  The generator state is part of the streaming state.  NULL switches back to injected uniforms. */
cpq_status cpq_set_dither_seed(cpq_handle h, const uint64_t* stream_seeds);
/* The same for uniforms that already live on the handle's device (16-byte aligned, same layout): borrowed, not copied -- the
 * buffer must stay valid until the process calls that use it have returned. */
cpq_status cpq_set_dither_uniforms_device(cpq_handle h, const double* d_uniforms, int64_t samples_per_channel);

/* convo::OutputFilter::prepare(sr) + process(block, convIsLast, hcMode, lcMode, lpMode) (OutputFilter.h:108-131,
 * OutputFilter.cpp:72-112,139-421): three cascaded DF2T biquads between the EQ and the makeup gain
 * (AudioEngine.Processing.DSPCoreDouble.cpp:452-463).  conv_is_last: low cut (lc_mode) + high cut (hc_mode, two
 * stages); otherwise 20 Hz high-pass + low-pass (lp_mode, two stages).  Runs when CPQ_STAGE_OUTPUT_FILTER is requested. */
cpq_status cpq_set_output_filter(cpq_handle h, int enabled, int conv_is_last, int hc_mode, int lc_mode, int lp_mode);
/* The rest of processOutputDouble around the headroom / dither step, inside CPQ_STAGE_EPILOGUE:
 * dc_cutoff_hz > 0: the output UltraHighRateDCBlocker pair (UltraHighRateDCBlocker.h:60-126; the engine uses 3.0 Hz,
 * AudioEngine.h:643-651) after the makeup gain; hard_clamp: the 1e300 / non-finite scrub and the
 * +-kOutputHeadroom clamp (DSPCoreDouble.cpp:665-691, 712-737).  SimplePeakLimiter (:700-710) sits between the two: see
 * cpq_set_peak_limiter. */
cpq_status cpq_set_output_stage(cpq_handle h, double dc_cutoff_hz, int hard_clamp);
/* Gain of the input stage (CPQ_STAGE_INPUT); applied when it differs from 1 by more than 1e-9, like the reference. */
cpq_status cpq_set_input_gain(cpq_handle h, double gain);
/* SimplePeakLimiter::prepare(sr, release_ms) + processBlock with the engine's constants (audioengine/SimplePeakLimiter.h,
 * DSPCoreDouble.cpp:700-710, threshold kOutputHeadroom - 0.5 dB, knee 1 dB; the engine prepares it with 100 ms,
 * AudioEngine.Processing.DSPCoreLifecycle.cpp:228), between the scrub and the hard clamp inside CPQ_STAGE_EPILOGUE; one
 * envelope per stream for both channels, reset to 1 at the start of every call.  release_ms = 0 (default) leaves the stage out. */
cpq_status cpq_set_peak_limiter(cpq_handle h, double release_ms);
/* state.convolverInputTrimGain: applied between the EQ and the convolver in the EQThenConvolver order when it differs from
 * 1 by more than 1e-12 (DSPCoreDouble.cpp:438-445). */
cpq_status cpq_set_conv_input_trim(cpq_handle h, double gain);
/* ConvolverProcessor::process with its smoothers settled (ConvolverProcessor.Runtime.cpp:367-377, 551-588, 675-677, 748;
 * cfg.conv_boundary = CPQ_CONV_OUTER): out = scrub(wet) * equalPowerSin(mix) + dry * equalPowerSin(1 - mix), dry = the
 * convolver's input delayed by dry_delay_samples (the reference's latency compensation: min(getLatency(), MAX_BLOCK_SIZE) +
 * min(irLatency, MAX_IR_LATENCY), see cpq_latency and cpq_ir_peak_latency); mix >= 0.999 drops the dry path, mix <= 0.001 is the
 * dry-only fast path (delayed input, no gain, no convolution).  mix is the float mixTarget.  Default 1.0 / 0. */
cpq_status cpq_set_mix(cpq_handle h, float mix, int dry_delay_samples);
/* ConvolverProcessor::setBypass as seen by process (runtimeSnapshot.bypassed -> processBypassWithLatencyCompensation,
 * ConvolverProcessor.Runtime.cpp:123-186, 257-261): the convolver stage outputs its input delayed by the dry_delay_samples of
 * cpq_set_mix (the reference uses latency + irLatency, latency = 0 with the direct head), whatever the mix.  CPQ_CONV_OUTER only. */
cpq_status cpq_set_convolver_bypass(cpq_handle h, int bypassed);
/* Host-only: LoaderThread::estimatePeakLatencySamples (convolver/ConvolverProcessor.LoaderThread.cpp:149-207), the irLatency
 * StereoConvolver::init receives as peakDelay: energy centroid of the first 99.9 % of the energy, maximum over channels
 * (ir_r nullable), rounded, clamped to [0, len - 1]. */
int cpq_ir_peak_latency(const double* ir_l, const double* ir_r, int len);
/* Host-only: IRConverter::computeScaleFactor (IRConverter.cpp:13-196; ScaleFactorResult IRConverter.h) -- the `scale`
 * StereoConvolver::init passes to SetImpulse: loudest channel to unit energy with a -6 dB margin, then the peak (0.5), RMS
 * (0.25) and frequency-response (+3 dB, IRAnalyzer::estimateMaxFrequencyResponseGain, IRAnalyzer.cpp:63-155) clamps, then the
 * jump protection against the IR that is currently playing (cur_* nullable / 0).  ir_r / cur_r nullable (mono). */
typedef struct cpq_ir_scale
{
    double scale_factor;
    int32_t has_scale_factor;
    float additional_attenuation_db;
} cpq_ir_scale;
cpq_status cpq_ir_scale_factor(const double* ir_l, const double* ir_r, int len, const double* cur_l, const double* cur_r, int cur_len,
                               double cur_scale, cpq_ir_scale* out);
/* Host-only: convertToMinimumPhase (convolver/ConvolverProcessor.ResampleAndFallback.cpp:333-460) for one channel:
 * homomorphic minimum-phase reconstruction (log magnitude -> real cepstrum folded onto its causal half -> exp) on an FFT of
 * nextPow2(4 len) points, written to out[len].  CPQ_ERR_UNSUPPORTED where the reference gives up and keeps the
 * linear-phase IR (FFT above 2^23 points, non-finite spectrum). */
cpq_status cpq_ir_min_phase(const double* ir, int len, double* out);
/* Host-only: what LoaderThread::doLoadStep does to one channel of a loaded IR (already at the device rate) before it reaches
 * SetImpulse (convolver/ConvolverProcessor.LoaderThread.cpp:588-637): 1 Hz UltraHighRateDCBlocker, asymmetric Tukey window
 * around the peak (ConvolverProcessor.ResampleAndFallback.cpp:111-196), copy into targetLength = min(int(sr * target_seconds),
 * 2^21) samples (computeTargetIRLength, ConvolverProcessor.StateAndUI.cpp:942-957; zero padded) with a linear fade-out over the
 * last 2 % (256 samples .. 80 ms) of the copied samples.  Returns the number of samples written (= cpq_ir_target_length), or -1
 * on a bad argument / out_capacity too small.  Sample-rate conversion is r8brain-free-src in the reference (third party) and
 * is not part of this library: resample first. */
int cpq_ir_target_length(double sample_rate, double target_seconds);
int cpq_ir_prepare(const double* ir, int len, double sample_rate, double target_seconds, double* out, int out_capacity);
/* Host-only: the IR file decode of LoaderThread::doLoadStep (convolver/ConvolverProcessor.LoaderThread.cpp:439-486).  The
 * reference reads through JUCE's WavAudioFormat (external, not in the reference tree: restated, parity unpinned): 32-bit float
 * samples -- integer PCM left-justified to 32 bits times 1.0f / 0x7fffffff, IEEE samples narrowed to float -- widened to double
 * and passed through the input transform (NaN / |v| < 1e-20 -> 0, clamp to [-1, 1], InputBitDepthTransform.h:86-121).  RIFF/WAVE
 * PCM 8/16/24/32 and IEEE float 32/64 (also as WAVE_FORMAT_EXTENSIBLE); anything else is CPQ_ERR_UNSUPPORTED.  `out` (nullable)
 * receives [channels][frames] doubles; call once with out = NULL for the sizes. */
typedef struct cpq_ir_file
{
    int32_t channels;
    int32_t bits_per_sample;
    int32_t is_float;
    int64_t frames;
    double sample_rate;
} cpq_ir_file;
cpq_status cpq_ir_decode_wav(const void* bytes, size_t n, cpq_ir_file* info, double* out, size_t out_capacity);
/* Host-only: doTrimStep's trailing-silence trim (:496-547): samples up to the last one above 1e-15 in either channel (ch1
 * nullable), at least 1. */
int cpq_ir_trim_silence(const double* ch0, const double* ch1, int n);
/* Host-only: convertToMixedPhaseFallback (convolver/ConvolverProcessor.MixedPhase.cpp:721-865), one channel: the linear-phase
 * IR's magnitude with the minimum-phase IR's phase below lo_hz, the delay of the linear IR's peak above hi_hz, raised-cosine blend
 * between (reference defaults 200 / 1000 Hz).  The reference's first choice, an optimised all-pass cascade with a disk cache
 * (:68-719), is a design tool and is not part of this library. */
cpq_status cpq_ir_mixed_phase(const double* linear, const double* minimum, int len, double sample_rate, double lo_hz, double hi_hz, double* out);
/* The loader pipeline for one stream (LoaderThread::doLoadStep / doTrimStep / doTransformStep / doBuildStep, :430-757): decode,
 * trailing-silence trim, cpq_ir_prepare per channel (target_seconds: the reference's default is 1.0), phase_mode 0 as is /
 * 1 minimum phase / 2 mixed phase (fallback form; either is skipped like in the reference when its result is not finite or
 * silent), cpq_ir_scale_factor, cpq_ir_peak_latency, then SetImpulse for every channel of the stream (a mono file feeds both).
 * A file at another sample rate is CPQ_ERR_UNSUPPORTED: the reference resamples with r8brain-free-src (third party). */
typedef struct cpq_ir_load_info
{
    int32_t file_channels, file_frames;
    int32_t trimmed_frames, target_length;
    int32_t peak_latency;        /* irLatency: the dry-path delay of cpq_set_mix */
    int32_t phase_applied;       /* 0 / 1 / 2: what the IR went through */
    double file_sample_rate;
    double scale_factor;
} cpq_ir_load_info;
cpq_status cpq_load_impulse_wav(cpq_handle h, int stream, const void* bytes, size_t n, int phase_mode, double target_seconds,
                                const cpq_filter_spec* spec, cpq_ir_load_info* info);
/* Host-only: IRAnalyzer::estimateMaxFrequencyResponseGain on its own (linear gain; 1.0 for an empty IR). */
double cpq_ir_freq_peak_gain(const double* ir_l, const double* ir_r, int len);
/* Host-only: the three stages' normalised coefficients {b0,b1,b2,a1,a2} x 3 as OutputFilter::prepare computes them. */
void cpq_output_filter_design(double sample_rate, int conv_is_last, int hc_mode, int lc_mode, int lp_mode, double out[15]);

/* convo::EQBandParams (core/EQParameters.h:11-20) as plain C. */
typedef struct cpq_eq_band_params
{
    float frequency, gain_db, q;
    int32_t enabled, type, channel_mode;   /* type: 0 LowShelf 1 Peaking 2 HighShelf 3 LowPass 4 HighPass */
} cpq_eq_band_params;

/* Host-only: EQProcessor::loadFromTextFile (eqprocessor/EQProcessor.Core.cpp:300-495), the EqualizerAPO / AutoEq
 * "ParametricEq.txt" format ("Preamp: -6.5 dB", "Channel: L R", "Filter 1: ON PK Fc 100 Hz Gain -3 dB Q 1.41").  `text` is
 * the file's contents.  bands / total_gain_db are in/out like the processor's state: every band is first disabled, set to
 * Stereo and 0 dB; frequency, Q and type of bands the text does not mention keep what the caller put there.  Returns the
 * number of "Filter" lines beyond band 20 that were ignored (the reference shows a warning box), or -1 on a NULL argument. */
int cpq_parse_eq_preset(const char* text, cpq_eq_band_params bands[CPQ_NUM_BANDS], float* total_gain_db);

/* EQProcessor::calcSVFCoeffs (eqprocessor/EQProcessor.Coefficients.cpp:101-130,431-618), host-side:
 * float parameters clamped then promoted to double exactly like the reference.
 * type: 0 LowShelf 1 Peaking 2 HighShelf 3 LowPass 4 HighPass. */
cpq_status cpq_design_band(int type, float freq_hz, float gain_db, float q, double sample_rate, cpq_svf_coeffs* out);
double cpq_db_to_gain(float db);            /* juce::Decibels::decibelsToGain<double>((double)db) */
double cpq_equal_power_sin(double x);       /* ConvolverProcessor.Runtime.cpp:26-31 */

/* ---- process ----------------------------------------------------------------------------- */
/* In place, planar, host buffers: planar[stream*n_channels + ch] points at T doubles.  Equivalent to the
 * reference running T/block_size callbacks of DSPCore::processDouble's conv -> EQ -> output chain on each
 * stream from a reset state.  T must be a multiple of block_size and <= max_samples.  H2D and D2H copies
 * are part of the call. */
cpq_status cpq_process(cpq_handle h, double* const* planar, int64_t T, unsigned stages);
/* Host buffers may be pinned (cudaHostAlloc / cudaHostRegister: the copies run at the PCIe rate) or ordinary pageable memory
 * (rows go through pinned staging slots filled and drained by a few host threads, CPQ_STAGE_THREADS in the environment sets
 * their number, 0 = leave it to the driver).  T may be odd (441-sample hosts).
 * Non-finite or enormous samples: where the reference zeroes a band state that reached 1e15 or went non-finite and carries on
 * (EQProcessor.Processing.cpp:174-175, :257-258) so does the engine -- the affected 1024-sample segment of that band runs with
 * the literal per-sample recurrence.
 * Error exits: no copy is in flight when a process call returns, whatever the status.  The call works in place, so after an
 * error the buffers hold a mixture of input and results. */

/* Same with FP32 host buffers, in place: the wire format of hosts that hand the application float blocks (its float path casts
 * on entry, convertFloatToDoubleHighQuality InputBitDepthTransform.h:102-121, and on exit, static_cast<float>,
 * AudioEngine.Processing.DSPCoreIO.cpp:524-537).  Samples are promoted to FP64 on the device, every stage computes in FP64 exactly
 * as in cpq_process, and the result is rounded to float on the way out; half the PCIe traffic.  Add CPQ_STAGE_INPUT for the
 * input transform that follows the cast in the reference. */
cpq_status cpq_process_f32(cpq_handle h, float* const* planar, int64_t T, unsigned stages);

/* Same, data already resident: d_io is a device pointer to [n_streams*n_channels][stride] doubles; stride even (rows are
 * 16-byte aligned) and >= T, i.e. >= T + 1 when T is odd (441-sample hosts): the pad sample of a row may be overwritten. */
cpq_status cpq_process_device(cpq_handle h, double* d_io, int64_t stride, int64_t T, unsigned stages);

/* Partition-range sharding for very long IRs (cfg 5): only layers/partitions in
 * [part_begin, part_end) of the flattened (layer, partition) list contribute to the convolver sum.
 * cpq_process_device(..., CPQ_STAGE_CONV) then yields this rank's partial y; the caller reduces the
 * partials (NCCL) and runs the remaining stages with cpq_process_device(..., CPQ_STAGE_EQ|EPILOGUE). */
cpq_status cpq_set_partition_range(cpq_handle h, int part_begin, int part_end);
int cpq_total_partitions(cpq_handle h);
/* The reduce step without a separate collective: device pointers (valid in this process: peer memory mapped over NVLink,
 * e.g. torch symmetric memory or CUDA IPC) of every rank's partial buffer, this rank's own included, in rank order and in
 * the layout / stride of the d_io passed to cpq_process_device.  While set, a cpq_process_device call WITHOUT CPQ_STAGE_CONV
 * sums the n buffers in that order as it loads its tiles (the same order on every rank, so all ranks see identical sums)
 * and writes the result into d_io.  The caller orders the ranks (a barrier after every rank's CPQ_STAGE_CONV call, and
 * before the buffers are reused).  n = 0 clears.  At most CPQ_MAX_PEERS (8). */
#define CPQ_MAX_PEERS 8
cpq_status cpq_set_partial_sources(cpq_handle h, int n, const double* const* device_ptrs);
/* Restrict the following cpq_process / cpq_process_device calls to streams [first_stream, first_stream + n_streams)
 * (n_streams = -1: through the last stream): with partition-range sharding every rank convolves all streams but finishes
 * (EQ, output stages) only the streams it owns.  Buffers keep their full layout. */
cpq_status cpq_set_stream_window(cpq_handle h, int first_stream, int n_streams);

/* ---- introspection ----------------------------------------------------------------------- */
cpq_status cpq_get_layout(cpq_handle h, cpq_layout* out);
int cpq_latency(cpq_handle h);                       /* getLatency(), MKLNonUniformConvolver.cpp:1055 */
cpq_status cpq_get_timings(cpq_handle h, cpq_timings* out);   /* CUDA-event times of the last cpq_process */
cpq_status cpq_get_eq_state(cpq_handle h, int stream, double* out /* [n_channels][20][2] ic1eq, ic2eq */);
void* cpq_cuda_stream(cpq_handle h);                 /* cudaStream_t the kernels are launched on */
int64_t cpq_kernel_launch_count(cpq_handle h);       /* launches since create */

/* Debugging aid: with CPQ_GUARD=1 in the environment every device buffer the library allocates carries a 256-byte canary on
 * either side; returns the number of canaries that have been overwritten (0 = clean), -1 when guard mode is off. */
int cpq_debug_check_guards(void);

/* Diagnostics (bench.py's FP64 roofline denominator): measured DFMA throughput in TFLOP/s (2 flops per DFMA) of `device`
 * over `iters` dependent-chain iterations per thread, and the dependent-issue latency of one DFMA in cycles; < 0 on error. */
double cpq_probe_dfma_tflops(int device, int iters);
double cpq_probe_dfma_latency(int device);

/* Host-only: the layer plan + per-callback gather plan without a device (used by tests and by
 * INTEGRATION.md's binding to validate geometry). src_offsets (nullable) receives, for each layer >= 1 and
 * each of n_callbacks callbacks, the delay-line stream position read (or -1 = skipped), layer-major. */
cpq_status cpq_plan_layout(int ir_len, int block_size, const cpq_filter_spec* spec, int64_t n_callbacks,
                           cpq_layout* out, int64_t* src_offsets);

/* The same for a host whose block is not a power of two: SetImpulse(known_block_size = the block rounded up to a power of
 * two), Add/Get calls of call_size samples.  l0_src / l0_count (nullable, n_callbacks each): the position in the L0 output
 * stream and the number of samples the reference's output ring delivers to each callback (the rest of the callback is zero). */
cpq_status cpq_plan_layout_ex(int ir_len, int known_block_size, int call_size, const cpq_filter_spec* spec, int64_t n_callbacks,
                              cpq_layout* out, int64_t* src_offsets, int64_t* l0_src, int32_t* l0_count);

#ifdef __cplusplus
}
#endif
#endif /* CPQ_H */
