/*
 * cpq_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of ConvoPeq's DSP hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (convopeq_b200/, include/) never does.
 *
 * Parity status: PINNED.  The reference holds no golden outputs for this path,
 * so the pin is the reference itself: oracle/_ref/libcpq_ref.so is the reference's own translation
 * units compiled in place (oracle/Makefile) and tests/test_oracle_vs_ref.py checks this file against
 * it whenever it is present; tests/golden/ holds outputs of that library (generator:
 * tests/golden/make_golden.py) which tests/test_oracle_golden.py checks everywhere else.
 * PARITY UNPINNED for the pieces whose reference files need the whole JUCE application to compile and therefore cannot be
 * run here: cpqo_outer_mix (ConvolverProcessor::process' dry/wet mix), cpqo_ir_peak_latency / the arithmetic of
 * cpqo_ir_scale_factor around its pinned FFT stage (LoaderThread / IRConverter), the Tukey window and trim / fade of
 * cpqo_ir_prepare (its DC-blocker stage is pinned), cpqo_parse-free preset parsing lives in the product and is checked on the
 * reference's own fixture only, the product's WAV decode restates JUCE's reader (not in the reference tree; checked against an
 * independent decoder on the reference's sample files) and its mixed-phase fallback form is checked against numpy, and the
 * uniform-partition extension flag (not a reference mode at all).  The dither branch
 * (cpqo_epilogue_ex) IS pinned, bit for bit, against PsychoacousticDither.h compiled in place.  Each says so at its definition; they restate the cited source lines only.
 *
 * Written from the reference's behaviour, one callback at a time, deliberately in the reference's
 * own real-time formulation (ring buffers, FDL, time-sliced tail MAC) so that it is an independent
 * check of the product's offline/batched formulation.  Every function cites the reference lines
 * (paths relative to the reference's src/) it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CPQO_PI 3.14159265358979323846

/* ------------------------------------------------------------------------------------------
 * FFT.  The reference delegates to Intel IPP (FFTBackend.cpp:123-151): forward real -> CCS
 * [re0,im0,...,re(N/2),im(N/2)] unscaled, inverse CCS -> real scaled by 1/N
 * (IPP_FFT_DIV_INV_BY_N).  IPP is an external, unvendored dependency (found via IPPROOT, oneAPI
 * 2026.0 per README.md:24); its published contract is restated with a plain radix-2 FFT.
 * ------------------------------------------------------------------------------------------ */
typedef struct
{
    int n;       /* real length, power of two */
    int log2n;
    double* wr;  /* cos(2 pi k / n), k < n/2 */
    double* wi;  /* -sin(2 pi k / n) */
    int* rev;    /* bit reversal for n points */
    double* tr;  /* scratch */
    double* ti;
} cpqo_fft;

static int cpqo_ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

static int cpqo_next_pow2(int n)
{
    /* juce::nextPowerOfTwo semantics: smallest power of two >= n (n >= 1). */
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

static cpqo_fft* cpqo_fft_create(int n)
{
    cpqo_fft* f = (cpqo_fft*) calloc(1, sizeof(cpqo_fft));
    f->n = n;
    f->log2n = cpqo_ilog2(n);
    f->wr = (double*) malloc(sizeof(double) * (size_t) (n / 2 + 1));
    f->wi = (double*) malloc(sizeof(double) * (size_t) (n / 2 + 1));
    f->rev = (int*) malloc(sizeof(int) * (size_t) n);
    f->tr = (double*) malloc(sizeof(double) * (size_t) n);
    f->ti = (double*) malloc(sizeof(double) * (size_t) n);
    for (int k = 0; k < n / 2; ++k)
    {
        /* octant-reduced angles keep every twiddle correctly rounded to ~1 ulp */
        const double a = 2.0 * CPQO_PI * (double) k / (double) n;
        f->wr[k] = cos(a);
        f->wi[k] = -sin(a);
    }
    for (int i = 0; i < n; ++i)
    {
        int r = 0;
        for (int b = 0; b < f->log2n; ++b)
            if (i & (1 << b)) r |= 1 << (f->log2n - 1 - b);
        f->rev[i] = r;
    }
    return f;
}

static void cpqo_fft_destroy(cpqo_fft* f)
{
    if (!f) return;
    free(f->wr); free(f->wi); free(f->rev); free(f->tr); free(f->ti);
    free(f);
}

/* in-place radix-2 DIT on (tr, ti), input already bit-reversed; sign = -1 forward, +1 inverse */
static void cpqo_fft_core(const cpqo_fft* f, double* tr, double* ti, int inverse)
{
    const int n = f->n;
    for (int len = 2; len <= n; len <<= 1)
    {
        const int half = len >> 1;
        const int step = n / len;
        for (int i = 0; i < n; i += len)
        {
            for (int j = 0; j < half; ++j)
            {
                const double wr = f->wr[j * step];
                const double wi = inverse ? -f->wi[j * step] : f->wi[j * step];
                const int a = i + j, b = a + half;
                const double xr = tr[b] * wr - ti[b] * wi;
                const double xi = tr[b] * wi + ti[b] * wr;
                tr[b] = tr[a] - xr;
                ti[b] = ti[a] - xi;
                tr[a] += xr;
                ti[a] += xi;
            }
        }
    }
}

/* forward real -> CCS (n+2 doubles). FFTBackend.cpp:123-135 contract. */
static void cpqo_fft_fwd(cpqo_fft* f, const double* in, double* ccs)
{
    const int n = f->n;
    for (int i = 0; i < n; ++i)
    {
        f->tr[f->rev[i]] = in[i];
        f->ti[f->rev[i]] = 0.0;
    }
    cpqo_fft_core(f, f->tr, f->ti, 0);
    for (int k = 0; k <= n / 2; ++k)
    {
        ccs[2 * k] = f->tr[k];
        ccs[2 * k + 1] = f->ti[k];
    }
    ccs[1] = 0.0;
    ccs[n + 1] = 0.0;
}

/* inverse CCS -> real, scaled by 1/n. FFTBackend.cpp:139-151 contract (imag of bins 0, n/2 ignored). */
static void cpqo_fft_inv(cpqo_fft* f, const double* ccs, double* out)
{
    const int n = f->n;
    for (int k = 0; k <= n / 2; ++k)
    {
        double re = ccs[2 * k], im = ccs[2 * k + 1];
        if (k == 0 || k == n / 2) im = 0.0;
        f->tr[f->rev[k]] = re;
        f->ti[f->rev[k]] = im;
        if (k > 0 && k < n / 2)
        {
            f->tr[f->rev[n - k]] = re;
            f->ti[f->rev[n - k]] = -im;
        }
    }
    cpqo_fft_core(f, f->tr, f->ti, 1);
    const double s = 1.0 / (double) n;
    for (int i = 0; i < n; ++i) out[i] = f->tr[i] * s;
}

/* ------------------------------------------------------------------------------------------
 * Non-uniform partitioned convolver, one channel.  MKLNonUniformConvolver.{h,cpp}.
 * ------------------------------------------------------------------------------------------ */
typedef struct
{
    double sample_rate;
    int hc_mode;   /* 0 Sharp, 1 Natural, 2 Soft   (OutputFilter.h:75-80) */
    int lc_mode;   /* 0 Natural, 1 Soft            (OutputFilter.h:85-89) */
    int tail_mode; /* 0 air absorption, 1 layer tail contouring, 2 bypass (MKLNonUniformConvolver.h:128) */
    int tail_enabled;
    double tail_start_seconds;
    double tail_strength;
    int tail_l1l2_multiplier;
} cpqo_filter_spec;

typedef struct
{
    int fft_size, part_size, num_parts, num_parts_ir, fdl_mask, complex_size;
    int is_immediate;
    double* ir_re; /* [num_parts][complex_size], reference order (reversed partitions) */
    double* ir_im;
    double* fdl_re; /* [2*num_parts][complex_size], mirrored ring */
    double* fdl_im;
    double* fft_time;  /* [fft_size] */
    double* fft_out;   /* [fft_size] */
    double* prev_in;   /* [part_size] */
    double* in_acc;    /* [part_size] */
    double* acc_re;    /* [complex_size] */
    double* acc_im;
    double* ccs;       /* [fft_size+2] scratch */
    int fdl_index, input_pos;
    double* tail_out;  /* [part_size] */
    int output_delay_samples, delay_cap;
    double* delay_buf;
    uint64_t delay_w, delay_r;
    int parts_per_callback, next_part, base_fdl_idx_saved, distributing;
    cpqo_fft* fft;
} cpqo_layer;

typedef struct
{
    cpqo_layer layers[3];
    int num_layers;
    int latency;
    double* ring;
    int ring_size, ring_mask, ring_w, ring_r, ring_avail;
    int ready, tail_enabled, max_block;
    double layer_gain[3];
    /* direct-form head (enableDirectHead, :689-718, 1169-1232, 1604-1616) */
    int direct_taps, direct_pending;
    double direct_ir_rev[32], direct_hist[32];
    double* direct_out;
} cpqo_nuc;

static double cpqo_clampd(double lo, double hi, double v) { return v < lo ? lo : (hi < v ? hi : v); }
static int cpqo_clampi(int lo, int hi, int v) { return v < lo ? lo : (hi < v ? hi : v); }

static void cpqo_layer_free(cpqo_layer* l)
{
    free(l->ir_re); free(l->ir_im); free(l->fdl_re); free(l->fdl_im); free(l->fft_time); free(l->fft_out);
    free(l->prev_in); free(l->in_acc); free(l->acc_re); free(l->acc_im); free(l->ccs); free(l->tail_out);
    free(l->delay_buf);
    cpqo_fft_destroy(l->fft);
    memset(l, 0, sizeof(*l));
}

static void cpqo_nuc_release(cpqo_nuc* c)
{
    for (int i = 0; i < 3; ++i) cpqo_layer_free(&c->layers[i]);
    free(c->ring);
    c->ring = NULL;
    c->num_layers = 0;
    c->ready = 0;
}

cpqo_nuc* cpqo_nuc_create(void) { return (cpqo_nuc*) calloc(1, sizeof(cpqo_nuc)); }

void cpqo_nuc_destroy(cpqo_nuc* c)
{
    if (!c) return;
    cpqo_nuc_release(c);
    free(c->direct_out);
    free(c);
}

/* applySpectrumFilter, MKLNonUniformConvolver.cpp:336-443 */
static void cpqo_apply_spectrum_filter(cpqo_nuc* c, const cpqo_filter_spec* spec)
{
    const double fs = spec->sample_rate;
    const double nyquist = fs * 0.5;
    const double hc_start = (fs <= 48000.0) ? 18000.0 : 22000.0;
    const double hc_end = nyquist;
    const double lc_end = (spec->lc_mode == 1) ? 6.0 : 8.0;
    const double lc_start = (spec->lc_mode == 1) ? 15.0 : 18.0;

    for (int li = 0; li < c->num_layers; ++li)
    {
        cpqo_layer* l = &c->layers[li];
        const int N = l->fft_size, halfN = N / 2, cs = l->complex_size;
        double* gain = (double*) malloc(sizeof(double) * (size_t) cs);
        for (int k = 0; k < cs; ++k) gain[k] = 1.0;
        {
            const int k_start = (int) round(hc_start * N / fs);
            int k_end = (int) round(hc_end * N / fs);
            if (k_end > halfN) k_end = halfN;
            for (int k = 0; k < cs; ++k)
            {
                if (k <= k_start) continue;
                if (k <= k_end)
                {
                    const double denom = (double) (k_end - k_start);
                    const double x = (double) (k - k_start) / denom;
                    switch (spec->hc_mode)
                    {
                        case 0: gain[k] = 1.0 / sqrt(1.0 + pow(x, 8.0)); break;
                        case 1: gain[k] = 0.5 * (1.0 + cos(CPQO_PI * x)); break;
                        case 2: gain[k] = exp(-4.60517 * x * x); break;
                        default: break;
                    }
                }
            }
        }
        {
            const int k_end = (int) round(lc_end * N / fs);
            const int k_start = (int) round(lc_start * N / fs);
            for (int k = 0; k < cs; ++k)
            {
                if (k <= k_end) gain[k] = 0.0;
                else if (k < k_start)
                {
                    const int d = k_start - k_end;
                    const double denom = (double) (d > 1 ? d : 1);
                    const double x = (double) (k - k_end) / denom;
                    gain[k] *= 0.5 * (1.0 - cos(CPQO_PI * x));
                }
            }
        }
        for (int p = 0; p < l->num_parts; ++p)
        {
            double* re = l->ir_re + (size_t) p * cs;
            double* im = l->ir_im + (size_t) p * cs;
            for (int k = 0; k < cs; ++k) { re[k] *= gain[k]; im[k] *= gain[k]; }
        }
        free(gain);
    }
}

int cpqo_nuc_set_impulse_ex(cpqo_nuc* c, const double* impulse_in, int ir_len, int block_size, double scale, int direct_head,
                            const cpqo_filter_spec* fs);
int cpqo_nuc_set_impulse(cpqo_nuc* c, const double* impulse, int ir_len, int block_size, double scale,
                         const cpqo_filter_spec* fs)
{
    return cpqo_nuc_set_impulse_ex(c, impulse, ir_len, block_size, scale, 0, fs);
}

/* SetImpulse, MKLNonUniformConvolver.cpp:610-1149, including the experimental direct-form head (:689-718, :730-731): the
 * first min(irLen, 32) taps run as a direct FIR (scaled, NOT passed through the spectrum filter) and are zeroed in the
 * impulse the partitions are built from.  Returns 1 on success. */
int cpqo_nuc_set_impulse_ex(cpqo_nuc* c, const double* impulse_in, int ir_len, int block_size, double scale, int direct_head,
                            const cpqo_filter_spec* fs)
{
    c->ready = 0;
    if (!impulse_in || ir_len <= 0 || block_size <= 0) return 0;
    cpqo_nuc_release(c);
    free(c->direct_out);
    c->direct_out = NULL;
    c->direct_taps = 0;
    c->direct_pending = 0;
    double* impulse_copy = NULL;
    const double* impulse = impulse_in;
    if (direct_head & 1)
    {
        const int direct_part = cpqo_next_pow2(block_size > 64 ? block_size : 64);
        int taps = direct_part < 32 ? direct_part : 32;   /* kMaxDirectTaps */
        if (ir_len < taps) taps = ir_len;
        c->direct_taps = taps;
        memset(c->direct_hist, 0, sizeof(c->direct_hist));
        for (int i = 0; i < taps; ++i) c->direct_ir_rev[i] = impulse_in[taps - 1 - i] * scale;
        c->direct_out = (double*) calloc((size_t) (block_size > 1 ? block_size : 1), sizeof(double));
        impulse_copy = (double*) malloc(sizeof(double) * (size_t) ir_len);
        memcpy(impulse_copy, impulse_in, sizeof(double) * (size_t) ir_len);
        memset(impulse_copy, 0, sizeof(double) * (size_t) taps);
        impulse = impulse_copy;
    }

    /* tail-mode table, :626-684 */
    const int tail_mode = fs ? cpqo_clampi(0, 2, fs->tail_mode) : 1;
    const int tail_enabled = (tail_mode != 2) && (fs ? (fs->tail_enabled != 0) : 1);
    const double sr_tail = fs ? fs->sample_rate : 48000.0;
    double tail_start = fs ? cpqo_clampd(0.01, 0.80, fs->tail_start_seconds) : 0.085;
    const double user_strength = fs ? cpqo_clampd(0.0, 2.0, fs->tail_strength) : 1.0;
    double tail_strength = user_strength;
    int mult = fs ? cpqo_clampi(2, 16, fs->tail_l1l2_multiplier) : 8;
    double g1 = 1.0, g2 = 1.0;
    const double s01 = cpqo_clampd(0.0, 1.0, user_strength * 0.5);
    if (!tail_enabled) { tail_strength = 0.0; g1 = 0.0; g2 = 0.0; }
    else if (tail_mode == 0)
    {
        tail_start = cpqo_clampd(0.01, 0.80, tail_start > 0.055 ? tail_start : 0.055);
        mult = cpqo_clampi(2, 16, mult > 6 ? mult : 6);
        tail_strength = cpqo_clampd(0.0, 2.0, user_strength);
        g1 = cpqo_clampd(0.0, 2.0, tail_strength * (0.95 - 0.25 * s01));
        g2 = cpqo_clampd(0.0, 2.0, tail_strength * (0.80 - 0.45 * s01));
    }
    else if (tail_mode == 1)
    {
        tail_start = cpqo_clampd(0.01, 0.80, tail_start > 0.12 ? tail_start : 0.12);
        tail_strength = cpqo_clampd(0.0, 2.0, tail_strength > 1.25 ? tail_strength : 1.25);
        mult = cpqo_clampi(2, 16, mult > 8 ? mult : 8);
        g1 = cpqo_clampd(0.0, 2.0, tail_strength * (1.05 + 0.20 * s01));
        g2 = cpqo_clampd(0.0, 2.0, tail_strength * (0.82 + 0.12 * s01));
    }
    else { tail_strength = 0.0; g1 = 0.0; g2 = 0.0; }
    c->tail_enabled = tail_enabled;
    c->max_block = block_size;
    c->layer_gain[0] = 1.0; c->layer_gain[1] = g1; c->layer_gain[2] = g2;

    /* layer plan, :738-758 */
    const int l0_part = cpqo_next_pow2(block_size > 64 ? block_size : 64);
    const int l1_part = l0_part * mult;
    const int l2_part = l1_part * mult;
    const int l0_max = 32 * l0_part;
    const int l0_by_tail = (int) llround(tail_start * sr_tail);
    const int l0_target = cpqo_clampi(l0_part, l0_max, l0_by_tail);
    /* direct_head bit 1 = the uniform-partition extension of the product (not a reference mode): the whole IR in layer 0 */
    const int uniform_ext = (direct_head & 2) != 0;
    const int l0_len = uniform_ext ? ir_len : (ir_len < (tail_enabled ? l0_target : l0_max) ? ir_len : (tail_enabled ? l0_target : l0_max));
    int l1_len = 0, l2_len = 0;
    if (tail_enabled)
    {
        l1_len = ir_len - l0_len;
        if (l1_len > 64 * l1_part) l1_len = 64 * l1_part;
        if (l1_len < 0) l1_len = 0;
        l2_len = ir_len - l0_len - l1_len;
        if (l2_len < 0) l2_len = 0;
    }
    const int offs[3] = { 0, l0_len, l0_len + l1_len };
    const int lens[3] = { l0_len, l1_len, l2_len };
    const int parts[3] = { l0_part, l1_part, l2_part };

    int prev_total = 0;
    c->num_layers = 0;
    for (int li = 0; li < 3; ++li)
    {
        if (lens[li] <= 0) continue;
        cpqo_layer* l = &c->layers[c->num_layers];
        memset(l, 0, sizeof(*l));
        l->part_size = parts[li];
        l->fft_size = l->part_size * 2;
        l->is_immediate = (li == 0);
        l->complex_size = l->fft_size / 2 + 1;
        l->num_parts_ir = (lens[li] + l->part_size - 1) / l->part_size;
        l->num_parts = cpqo_next_pow2(l->num_parts_ir);
        l->fdl_mask = l->num_parts - 1;
        l->fft = cpqo_fft_create(l->fft_size);
        const size_t cs = (size_t) l->complex_size;
        l->ir_re = (double*) calloc((size_t) l->num_parts * cs, sizeof(double));
        l->ir_im = (double*) calloc((size_t) l->num_parts * cs, sizeof(double));
        l->fdl_re = (double*) calloc((size_t) l->num_parts * 2 * cs, sizeof(double));
        l->fdl_im = (double*) calloc((size_t) l->num_parts * 2 * cs, sizeof(double));
        l->fft_time = (double*) calloc((size_t) l->fft_size, sizeof(double));
        l->fft_out = (double*) calloc((size_t) l->fft_size, sizeof(double));
        l->prev_in = (double*) calloc((size_t) l->part_size, sizeof(double));
        l->in_acc = (double*) calloc((size_t) l->part_size, sizeof(double));
        l->acc_re = (double*) calloc(cs, sizeof(double));
        l->acc_im = (double*) calloc(cs, sizeof(double));
        l->ccs = (double*) calloc((size_t) l->fft_size + 2, sizeof(double));
        if (!l->is_immediate) l->tail_out = (double*) calloc((size_t) l->part_size, sizeof(double));

        /* partition spectra, :919-946 (scale applied only when |scale-1| > 1e-12) */
        const double* src = impulse + offs[li];
        for (int p = 0; p < l->num_parts; ++p)
        {
            memset(l->fft_time, 0, sizeof(double) * (size_t) l->fft_size);
            if (p < l->num_parts_ir)
            {
                const int start = p * l->part_size;
                int n = lens[li] - start;
                if (n > l->part_size) n = l->part_size;
                if (n > 0) memcpy(l->fft_time, src + start, sizeof(double) * (size_t) n);
            }
            cpqo_fft_fwd(l->fft, l->fft_time, l->ccs);
            if (fabs(scale - 1.0) > 1e-12)
                for (int k = 0; k < l->complex_size * 2; ++k) l->ccs[k] *= scale;
            for (int k = 0; k < l->complex_size; ++k)
            {
                l->ir_re[(size_t) p * cs + (size_t) k] = l->ccs[2 * k];
                l->ir_im[(size_t) p * cs + (size_t) k] = l->ccs[2 * k + 1];
            }
        }
        memset(l->fft_time, 0, sizeof(double) * (size_t) l->fft_size);
        /* reverse partition order, :959-985 */
        for (int pf = 0; pf < l->num_parts_ir / 2; ++pf)
        {
            const int pb = l->num_parts_ir - 1 - pf;
            for (int k = 0; k < l->complex_size; ++k)
            {
                double t = l->ir_re[(size_t) pf * cs + (size_t) k];
                l->ir_re[(size_t) pf * cs + (size_t) k] = l->ir_re[(size_t) pb * cs + (size_t) k];
                l->ir_re[(size_t) pb * cs + (size_t) k] = t;
                t = l->ir_im[(size_t) pf * cs + (size_t) k];
                l->ir_im[(size_t) pf * cs + (size_t) k] = l->ir_im[(size_t) pb * cs + (size_t) k];
                l->ir_im[(size_t) pb * cs + (size_t) k] = t;
            }
        }
        /* partsPerCallback, :988-994 */
        if (!l->is_immediate)
        {
            const int bs = block_size > 1 ? block_size : 1;
            const int blocks_per_part = (l->part_size + bs - 1) / bs;
            int ppc = (l->num_parts_ir + blocks_per_part - 1) / blocks_per_part;
            if (ppc < 1) ppc = 1;
            if (ppc > l->num_parts_ir) ppc = l->num_parts_ir;
            l->parts_per_callback = ppc;
        }
        /* delay line, :1004-1019 */
        if (prev_total > 0)
        {
            l->output_delay_samples = prev_total;
            l->delay_cap = ((prev_total + l->part_size + c->max_block + 15) / 16) * 16;
            l->delay_buf = (double*) calloc((size_t) l->delay_cap, sizeof(double));
        }
        ++c->num_layers;
        prev_total += lens[li];
    }
    if (c->num_layers == 0) { free(impulse_copy); return 0; }

    /* L0 output ring, :1033-1053 */
    {
        const int l0p = c->layers[0].part_size;
        const int npi = (ir_len + block_size - 1) / block_size;
        const int np = cpqo_next_pow2(npi);
        const int base = np * 2;
        const int margin = cpqo_next_pow2(block_size);
        const int rsize = cpqo_next_pow2(base + margin);
        const int minsize = cpqo_next_pow2(l0p * 4 + block_size * 4);
        const int fin = rsize > minsize ? rsize : minsize;
        c->ring = (double*) calloc((size_t) fin, sizeof(double));
        c->ring_size = fin;
        c->ring_mask = fin - 1;
        c->ring_w = c->ring_r = c->ring_avail = 0;
    }
    c->latency = c->layers[0].part_size;

    if (fs) cpqo_apply_spectrum_filter(c, fs);

    /* air-absorption tilt on L1/L2, :1060-1097 */
    if (tail_enabled && tail_mode == 0)
    {
        const double start_norm = cpqo_clampd(0.65, 1.55, tail_start / 0.085);
        const double damping_base = (0.35 + 1.10 * s01) * start_norm;
        for (int li = 1; li < c->num_layers; ++li)
        {
            cpqo_layer* l = &c->layers[li];
            const double w = (li == 1) ? 1.0 : 1.6;
            const double dc = damping_base * w;
            const int dn = l->complex_size - 1;
            const double denom = (double) (dn > 1 ? dn : 1);
            for (int k = 0; k < l->complex_size; ++k)
            {
                const double fn = (double) k / denom;
                const double tilt = exp(-dc * fn * fn);
                for (int p = 0; p < l->num_parts; ++p)
                {
                    l->ir_re[(size_t) p * l->complex_size + (size_t) k] *= tilt;
                    l->ir_im[(size_t) p * l->complex_size + (size_t) k] *= tilt;
                }
            }
        }
    }
    free(impulse_copy);
    c->ready = 1;
    return 1;
}

/* processDirectBlock, :1169-1232: y[n] = sum_k hrev[k] * window[n + k] with two 4-lane FMA accumulators over blocks of eight
 * taps, lanes summed as (v0 + v2) + (v1 + v3), remainder taps added one by one; |y| < 1e-20 or non-finite -> 0. */
static void cpqo_direct_block(cpqo_nuc* c, const double* input, int n)
{
    const int taps = c->direct_taps, hist = taps - 1;
    double* win = (double*) malloc(sizeof(double) * (size_t) (hist + n));
    memcpy(win, c->direct_hist, sizeof(double) * (size_t) hist);
    if (input) memcpy(win + hist, input, sizeof(double) * (size_t) n);
    else memset(win + hist, 0, sizeof(double) * (size_t) n);
    const int v8 = taps / 8 * 8;
    for (int i = 0; i < n; ++i)
    {
        const double* x = win + i;
        double s0[4] = { 0, 0, 0, 0 }, s1[4] = { 0, 0, 0, 0 };
        int k = 0;
        for (; k < v8; k += 8)
            for (int j = 0; j < 4; ++j)
            {
                s0[j] = fma(c->direct_ir_rev[k + j], x[k + j], s0[j]);
                s1[j] = fma(c->direct_ir_rev[k + 4 + j], x[k + 4 + j], s1[j]);
            }
        double v[4];
        for (int j = 0; j < 4; ++j) v[j] = s0[j] + s1[j];
        double y = (v[0] + v[2]) + (v[1] + v[3]);
        for (; k < taps; ++k) y += c->direct_ir_rev[k] * x[k];
        if (!(isfinite(y) && !(fabs(y) < 1.0e-20))) y = 0.0;
        c->direct_out[i] = y;
    }
    memcpy(c->direct_hist, win + n, sizeof(double) * (size_t) hist);
    free(win);
    c->direct_pending = n;
}

/* accumulateSplitComplex, :150-195 (separate multiplies and adds) */
static void cpqo_mac(const double* ar, const double* ai, const double* br, const double* bi, double* dr, double* di, int n)
{
    for (int k = 0; k < n; ++k)
    {
        dr[k] += ar[k] * br[k] - ai[k] * bi[k];
        di[k] += ar[k] * bi[k] + ai[k] * br[k];
    }
}

static void cpqo_ring_write(cpqo_nuc* c, const double* src, int n) /* :1341-1371 */
{
    if (n <= 0 || !c->ring) return;
    int first = c->ring_size - c->ring_w;
    if (first > n) first = n;
    memcpy(c->ring + c->ring_w, src, sizeof(double) * (size_t) first);
    if (n > first) memcpy(c->ring, src + first, sizeof(double) * (size_t) (n - first));
    c->ring_w = (c->ring_w + n) & c->ring_mask;
    const int next = c->ring_avail + n;
    if (next > c->ring_size)
    {
        c->ring_r = (c->ring_r + (next - c->ring_size)) & c->ring_mask;
        c->ring_avail = c->ring_size;
    }
    else c->ring_avail = next;
}

static int cpqo_ring_read(cpqo_nuc* c, double* dst, int n) /* :1376-1402 */
{
    if (n <= 0 || !c->ring) return 0;
    const int to_read = n < c->ring_avail ? n : c->ring_avail;
    if (to_read == 0)
    {
        if (dst) memset(dst, 0, sizeof(double) * (size_t) n);
        return 0;
    }
    int first = c->ring_size - c->ring_r;
    if (first > to_read) first = to_read;
    if (dst)
    {
        memcpy(dst, c->ring + c->ring_r, sizeof(double) * (size_t) first);
        if (to_read > first) memcpy(dst + first, c->ring, sizeof(double) * (size_t) (to_read - first));
        if (to_read < n) memset(dst + to_read, 0, sizeof(double) * (size_t) (n - to_read));
    }
    c->ring_r = (c->ring_r + to_read) & c->ring_mask;
    c->ring_avail -= to_read;
    return to_read;
}

/* frame assembly + forward FFT + FDL write with mirror, :1256-1283 / :1456-1479 */
static void cpqo_layer_push_frame(cpqo_layer* l)
{
    const size_t cs = (size_t) l->complex_size;
    memcpy(l->fft_time, l->prev_in, sizeof(double) * (size_t) l->part_size);
    memcpy(l->fft_time + l->part_size, l->in_acc, sizeof(double) * (size_t) l->part_size);
    memcpy(l->prev_in, l->in_acc, sizeof(double) * (size_t) l->part_size);
    cpqo_fft_fwd(l->fft, l->fft_time, l->ccs);
    double* r0 = l->fdl_re + (size_t) l->fdl_index * cs;
    double* i0 = l->fdl_im + (size_t) l->fdl_index * cs;
    double* r1 = l->fdl_re + (size_t) (l->fdl_index + l->num_parts) * cs;
    double* i1 = l->fdl_im + (size_t) (l->fdl_index + l->num_parts) * cs;
    for (int k = 0; k < l->complex_size; ++k)
    {
        r0[k] = r1[k] = l->ccs[2 * k];
        i0[k] = i1[k] = l->ccs[2 * k + 1];
    }
}

static void cpqo_layer_ifft(cpqo_layer* l)
{
    for (int k = 0; k < l->complex_size; ++k)
    {
        l->ccs[2 * k] = l->acc_re[k];
        l->ccs[2 * k + 1] = l->acc_im[k];
    }
    cpqo_fft_inv(l->fft, l->ccs, l->fft_out);
}

/* processLayerBlock (L0), :1245-1336 */
static void cpqo_process_layer_block(cpqo_nuc* c, cpqo_layer* l)
{
    const size_t cs = (size_t) l->complex_size;
    cpqo_layer_push_frame(l);
    memset(l->acc_re, 0, sizeof(double) * cs);
    memset(l->acc_im, 0, sizeof(double) * cs);
    const int lin_start = l->fdl_index - l->num_parts_ir + 1 + l->num_parts;
    for (int p = 0; p < l->num_parts_ir; ++p)
    {
        const size_t idx = (size_t) (lin_start + p);
        cpqo_mac(l->fdl_re + idx * cs, l->fdl_im + idx * cs, l->ir_re + (size_t) p * cs, l->ir_im + (size_t) p * cs,
                 l->acc_re, l->acc_im, l->complex_size);
    }
    cpqo_layer_ifft(l);
    cpqo_ring_write(c, l->fft_out + l->part_size, l->part_size);
    l->fdl_index = (l->fdl_index + 1) & l->fdl_mask;
}

static void cpqo_delay_write(cpqo_layer* l, const double* src, int n) /* :1639-1648 */
{
    const size_t off = (size_t) (l->delay_w % (uint64_t) l->delay_cap);
    int first = l->delay_cap - (int) off;
    if (first > n) first = n;
    memcpy(l->delay_buf + off, src, sizeof(double) * (size_t) first);
    if (first < n) memcpy(l->delay_buf, src + first, sizeof(double) * (size_t) (n - first));
    l->delay_w += (uint64_t) n;
}

static void cpqo_delay_read_add(cpqo_layer* l, double* dst, int n, double gain) /* :1653-1688 */
{
    if (!l->delay_buf || l->delay_cap <= 0 || !dst) return;
    const uint64_t max_read = (l->delay_w >= (uint64_t) l->output_delay_samples) ? l->delay_w - (uint64_t) l->output_delay_samples : 0;
    const uint64_t start = l->delay_r > max_read ? l->delay_r : max_read;
    if (start + (uint64_t) n > l->delay_w) return;
    const size_t off = (size_t) (start % (uint64_t) l->delay_cap);
    int first = l->delay_cap - (int) off;
    if (first > n) first = n;
    const int unity = fabs(gain - 1.0) < 1.0e-12;
    for (int i = 0; i < first; ++i) dst[i] += unity ? l->delay_buf[off + (size_t) i] : l->delay_buf[off + (size_t) i] * gain;
    for (int i = first; i < n; ++i) dst[i] += unity ? l->delay_buf[i - first] : l->delay_buf[i - first] * gain;
    l->delay_r = start + (uint64_t) n;
}

/* Add, :1407-1548 */
void cpqo_nuc_add(cpqo_nuc* c, const double* input, int n)
{
    if (!c->ready || n <= 0) return;
    if (c->direct_taps > 0)
    {
        if (n > c->max_block) c->direct_pending = 0;   /* :1174-1179: a call longer than the block drops the direct output */
        else cpqo_direct_block(c, input, n);
    }
    for (int li = 0; li < c->num_layers; ++li)
    {
        cpqo_layer* l = &c->layers[li];
        const size_t cs = (size_t) l->complex_size;
        int consumed = 0;
        while (consumed < n)
        {
            int fill = l->part_size - l->input_pos;
            if (fill > n - consumed) fill = n - consumed;
            if (input) memcpy(l->in_acc + l->input_pos, input + consumed, sizeof(double) * (size_t) fill);
            else memset(l->in_acc + l->input_pos, 0, sizeof(double) * (size_t) fill);
            l->input_pos += fill;
            consumed += fill;
            if (l->input_pos >= l->part_size)
            {
                l->input_pos = 0;
                if (l->is_immediate) cpqo_process_layer_block(c, l);
                else
                {
                    cpqo_layer_push_frame(l);
                    l->fdl_index = (l->fdl_index + 1) & l->fdl_mask;
                    l->base_fdl_idx_saved = (l->fdl_index - 1 + l->num_parts) & l->fdl_mask;
                    memset(l->acc_re, 0, sizeof(double) * cs);
                    memset(l->acc_im, 0, sizeof(double) * cs);
                    l->next_part = 0;
                    l->distributing = 1;
                }
            }
        }
        if (!l->is_immediate && l->distributing)
        {
            int end_part = l->next_part + l->parts_per_callback;
            if (end_part > l->num_parts_ir) end_part = l->num_parts_ir;
            const int lin_start = l->base_fdl_idx_saved - l->num_parts_ir + 1 + l->num_parts;
            for (int p = l->next_part; p < end_part; ++p)
            {
                const size_t idx = (size_t) (lin_start + p);
                cpqo_mac(l->fdl_re + idx * cs, l->fdl_im + idx * cs, l->ir_re + (size_t) p * cs, l->ir_im + (size_t) p * cs,
                         l->acc_re, l->acc_im, l->complex_size);
            }
            l->next_part = end_part;
            if (l->next_part >= l->num_parts_ir)
            {
                cpqo_layer_ifft(l);
                memcpy(l->tail_out, l->fft_out + l->part_size, sizeof(double) * (size_t) l->part_size);
                if (l->delay_buf) cpqo_delay_write(l, l->tail_out, l->part_size);
                l->distributing = 0;
                l->next_part = 0;
            }
        }
    }
}

/* Get, :1553-1634 */
int cpqo_nuc_get(cpqo_nuc* c, double* out, int n)
{
    if (!c->ready || n <= 0)
    {
        if (out && n > 0) memset(out, 0, sizeof(double) * (size_t) n);
        return 0;
    }
    const int got = cpqo_ring_read(c, out, n);
    if (c->direct_taps > 0)
    {
        const int to_add = n < c->direct_pending ? n : c->direct_pending;
        if (to_add > 0)
        {
            if (out)
                for (int i = 0; i < to_add; ++i) out[i] += c->direct_out[i];
            memset(c->direct_out, 0, sizeof(double) * (size_t) to_add);
            c->direct_pending = 0;
        }
    }
    for (int li = 1; li < c->num_layers; ++li)
    {
        cpqo_layer* l = &c->layers[li];
        if (!l->delay_buf || !out) continue;
        const double g = c->tail_enabled ? c->layer_gain[li] : 0.0;
        cpqo_delay_read_add(l, out, n, g);
    }
    return got;
}

/* StereoConvolver::process per channel (ConvolverProcessor.Runtime.cpp:1159-1184): Add, Get, zero-fill shortfall */
void cpqo_nuc_process(cpqo_nuc* c, const double* in, double* out, long total, int call)
{
    for (long pos = 0; pos < total; pos += call)
    {
        const int n = (int) ((total - pos) < call ? (total - pos) : call);
        cpqo_nuc_add(c, in ? in + pos : NULL, n);
        int got = cpqo_nuc_get(c, out + pos, n);
        if (got < 0) got = 0;
        if (got < n) memset(out + pos + got, 0, sizeof(double) * (size_t) (n - got));
    }
}

/* layout[l] = {partSize, numPartsIR, numParts, partsPerCallback, outputDelaySamples, delayLineCapacity, isImmediate, fftSize} */
int cpqo_nuc_layout(const cpqo_nuc* c, int* layout, double* gains)
{
    for (int l = 0; l < c->num_layers; ++l)
    {
        const cpqo_layer* L = &c->layers[l];
        int* o = layout + l * 8;
        o[0] = L->part_size; o[1] = L->num_parts_ir; o[2] = L->num_parts; o[3] = L->parts_per_callback;
        o[4] = L->output_delay_samples; o[5] = L->delay_cap; o[6] = L->is_immediate; o[7] = L->fft_size;
    }
    for (int l = 0; l < 3; ++l) gains[l] = c->layer_gain[l];
    return c->num_layers;
}

int cpqo_nuc_ir_spectrum(const cpqo_nuc* c, int layer, int part, double* re, double* im)
{
    if (layer < 0 || layer >= c->num_layers) return 0;
    const cpqo_layer* L = &c->layers[layer];
    if (part < 0 || part >= L->num_parts) return 0;
    memcpy(re, L->ir_re + (size_t) part * L->complex_size, sizeof(double) * (size_t) L->complex_size);
    memcpy(im, L->ir_im + (size_t) part * L->complex_size, sizeof(double) * (size_t) L->complex_size);
    return L->complex_size;
}

/* ConvolverProcessor::process at mix = 1 (Runtime.cpp:26-31,50-60,675,722,748): wet is scrubbed
 * (non-finite or |x| >= 1e300 -> 0) then multiplied by equalPowerSin(1.0) * CONVOLUTION_HEADROOM_GAIN(=1). */
double cpqo_equal_power_sin(double x)
{
    const double t = x * (CPQO_PI * 0.5);
    const double t2 = t * t;
    return t * (1.0 + t2 * (-1.0 / 6.0 + t2 * (1.0 / 120.0 + t2 * (-1.0 / 5040.0 + t2 * (1.0 / 362880.0)))));
}

void cpqo_outer_wet(double* data, long n, double mix)
{
    const double g = cpqo_equal_power_sin(mix) * 1.0;
    for (long i = 0; i < n; ++i)
    {
        double v = data[i];
        if (!(isfinite(v) && fabs(v) < 1.0e300)) v = 0.0;
        data[i] = v * g;
    }
}

/* The engine's input stage (convo::input_transform::applyHighQuality64BitTransform, InputBitDepthTransform.h:86-100, called
 * from DSPCore::processInput, AudioEngine.Processing.DSPCoreIO.cpp:203-232): optional gain, then sanitizeAndLimit (:32-68):
 * NaN or |v| < 1e-20 -> 0, clamp to [-1, 1].  The vector body (four samples at a time) keeps +-Inf and clamps it to +-1,
 * the scalar remainder (n % 4 samples) zeroes it; both are restated. */
void cpqo_input_transform(double* data, long n, double gain)
{
    const double gd = gain - 1.0;
    if (gd > 1e-9 || gd < -1e-9)
        for (long i = 0; i < n; ++i) data[i] *= gain;
    const long vend = n / 4 * 4;
    for (long i = 0; i < n; ++i)
    {
        double v = data[i];
        if (i < vend) { if (v != v || fabs(v) < 1.0e-20) v = 0.0; }
        else if (!(isfinite(v) && !(fabs(v) < 1.0e-20))) v = 0.0;
        v = v > -1.0 ? v : -1.0;   /* _mm256_max_pd(v, vMin) then min: applied after the zeroing, so NaN never reaches it */
        v = v < 1.0 ? v : 1.0;
        data[i] = v;
    }
}

/* ConvolverProcessor::process in its settled state (mix and latency smoothers at their targets, no bypass):
 * ConvolverProcessor.Runtime.cpp:367-377 (needsConvolution = mix > 0.001, needsDrySignal = mix < 0.999), :551-568 (dry =
 * input delayed by round(totalLatency) through a zero-initialised ring), :573-584 (dry-only fast path copies the dry signal,
 * no gain), :675-677 + :748 (out = scrub(wet) * equalPowerSin(mix) * CONVOLUTION_HEADROOM_GAIN + dry * equalPowerSin(1 - mix),
 * two products and one sum).  mix is the float mixTarget promoted to double.  PARITY UNPINNED: ConvolverProcessor needs
 * the JUCE application to compile, so this part restates the source and is not checked against the running reference. */
void cpqo_outer_mix(double* wet_io, const double* dry_in, long n, float mix, int delay)
{
    const double m = (double) mix;
    const int needs_conv = m > 0.001, needs_dry = m < 0.999;
    const double wg = cpqo_equal_power_sin(m) * 1.0;
    const double dg = needs_dry ? cpqo_equal_power_sin(1.0 - m) : 0.0;
    for (long i = 0; i < n; ++i)
    {
        const double dry = (i >= delay) ? dry_in[i - delay] : 0.0;
        if (!needs_conv) { wet_io[i] = dry; continue; }
        double v = wet_io[i];
        if (!(isfinite(v) && fabs(v) < 1.0e300)) v = 0.0;
        const double a = v * wg, b = dry * dg;
        wet_io[i] = a + b;
    }
}

/* LoaderThread::doLoadStep after the resampler (convolver/ConvolverProcessor.LoaderThread.cpp:588-637), one channel:
 * UltraHighRateDCBlocker at 1 Hz (UltraHighRateDCBlocker.h:78-188; PINNED against the header compiled in place), the asymmetric
 * Tukey window (ConvolverProcessor.ResampleAndFallback.cpp:111-196) and the trim to targetLength with its fade-out
 * (LoaderThread.cpp:619-637, ConvolverProcessor.StateAndUI.cpp:942-957) -- those two UNPINNED (their files need JUCE and
 * r8brain to compile).  Returns targetLength; out must hold that many samples. */
void cpqo_ir_dc_block(double* d, int n, double sr, double cutoff)
{
    double al[2] = { 1.0e-6, 1.0e-6 };
    if (isfinite(sr) && sr > 0.0 && isfinite(cutoff) && cutoff > 0.0)
        for (int i = 0; i < 2; ++i)
        {
            const double fc = cutoff * (i == 0 ? 1.0 - 0.1 : 1.0 + 0.1);
            double a = -expm1(-(2.0 * CPQO_PI * fc / sr));
            if (!isfinite(a) || a <= 0.0 || a >= 1.0) a = 1.0e-6;
            al[i] = a;
        }
    double s0 = 0.0, s1 = 0.0;
    for (int i = 0; i < n; ++i)
    {
        double x = d[i];
        s0 = fma(al[0], x - s0, s0);
        x = x - s0;
        s1 = fma(al[1], x - s1, s1);
        x = x - s1;
        d[i] = x;
    }
}

int cpqo_ir_prepare(const double* in, int len, double sr, double target_seconds, double* out)
{
    int target = (int) (sr * (double) (float) target_seconds);
    if (target > 2097152) target = 2097152;
    if (target < 1) target = 1;
    double* w = (double*) malloc(sizeof(double) * (size_t) (len > 0 ? len : 1));
    memcpy(w, in, sizeof(double) * (size_t) len);
    cpqo_ir_dc_block(w, len, sr, 1.0);
    /* asymmetric Tukey around the (first) largest |sample| */
    int peak = 0;
    for (int i = 1; i < len; ++i)
        if (fabs(w[i]) > fabs(w[peak])) peak = i;
    double a_post = 0.05 + 0.033 * (log2((double) len) - 10.0);
    if (a_post < 0.05) a_post = 0.05;
    if (a_post > 0.25) a_post = 0.25;
    if (peak > 0)
    {
        const int n = (int) floor(peak * 0.05);
        for (int i = 0; i < n; ++i) w[i] *= 0.5 * (1.0 + cos(CPQO_PI / (peak * 0.05) * (double) i - CPQO_PI));
    }
    const double dist = (double) (len - 1 - peak);
    if (dist > 1.0e-9)
    {
        const int start = peak + (int) ceil(dist * (1.0 - a_post));
        const double scale = (CPQO_PI / a_post) / dist;
        const double offset = (CPQO_PI / a_post) * (((double) start - (double) peak) / dist - (1.0 - a_post));
        for (int i = start; i < len; ++i) w[i] *= 0.5 * (1.0 + cos(scale * (double) (i - start) + offset));
    }
    memset(out, 0, sizeof(double) * (size_t) target);
    const int copy = target < len ? target : len;
    int fade = (int) round((double) copy * 0.02);
    int max_fade = (int) round(sr * 0.080);
    if (max_fade < 256) max_fade = 256;
    if (fade < 256) fade = 256;
    if (fade > max_fade) fade = max_fade;
    if (fade > copy - 1) fade = copy - 1;
    if (fade < 0) fade = 0;
    memcpy(out, w, sizeof(double) * (size_t) copy);
    double g = 1.0;
    for (int i = 0; i < fade; ++i)
    {
        out[copy - fade + i] *= g;
        g += (0.0 - 1.0) / (double) fade;
    }
    free(w);
    return target;
}

/* estimatePeakLatencySamples, convolver/ConvolverProcessor.LoaderThread.cpp:149-207: energy centroid of the first 99.9 % of
 * each channel's energy, maximum over channels, rounded half up and clamped to [0, len - 1].  This is the irLatency the
 * dry path is delayed by (on top of the algorithm latency). */
int cpqo_ir_peak_latency(const double* ir_l, const double* ir_r, int len)
{
    if (len <= 0) return 0;
    double max_centroid = 0.0;
    const double* chans[2] = { ir_l, ir_r };
    for (int c = 0; c < 2; ++c)
    {
        const double* d = chans[c];
        if (!d) continue;
        double total = 0.0;
        for (int i = 0; i < len; ++i) total += d[i] * d[i];
        if (total < 1e-12) continue;
        double cum = 0.0;
        int cutoff = len - 1;
        for (int i = 0; i < len; ++i)
        {
            cum += d[i] * d[i];
            if (cum >= total * 0.999) { cutoff = i; break; }
        }
        double se = 0.0, sw = 0.0;
        for (int i = 0; i <= cutoff; ++i)
        {
            const double e = d[i] * d[i];
            se += e;
            sw += (double) i * e;
        }
        const double centroid = se > 0.0 ? sw / se : 0.0;
        if (centroid > max_centroid) max_centroid = centroid;
    }
    int lat = (int) floor(max_centroid + 0.5);
    if (lat < 0) lat = 0;
    if (lat > len - 1) lat = len - 1;
    return lat;
}

/* IRAnalyzer::estimateMaxFrequencyResponseGain, IRAnalyzer.cpp:63-155: Tukey(0.5)-windowed radix-2 FFT of the first
 * min(len, 65536) samples (twiddles by repeated multiplication, like simpleRealFFT :14-59), largest bin magnitude over the
 * channels refined by the three-point log-parabolic interpolation, divided by the window's mean over the copied samples. */
double cpqo_ir_freq_peak_gain(const double* ir_l, const double* ir_r, int len)
{
    if (len <= 0 || !ir_l) return 1.0;
    const int copy_len = len < 65536 ? len : 65536;
    const int n = cpqo_next_pow2(copy_len);
    if (n < 2) return 1.0;
    const double pi = CPQO_PI, alpha = 0.5;
    const double taper = alpha * (double) (n - 1) * 0.5;
    double* win = (double*) malloc(sizeof(double) * (size_t) n);
    for (int i = 0; i < n; ++i)
    {
        const double t = (double) i;
        if (t < taper) win[i] = 0.5 * (1.0 + cos((2.0 * pi * t) / (alpha * (double) (n - 1)) - pi));
        else if (t > (double) (n - 1) - taper) win[i] = 0.5 * (1.0 + cos((2.0 * pi * (t - ((double) (n - 1) - taper))) / (alpha * (double) (n - 1))));
        else win[i] = 1.0;
    }
    double wsum = 0.0;
    for (int i = 0; i < copy_len; ++i) wsum += win[i];
    const double wmean = wsum / (double) copy_len;
    if (wmean < 1e-18) { free(win); return 1.0; }
    double* re = (double*) malloc(sizeof(double) * (size_t) n);
    double* im = (double*) malloc(sizeof(double) * (size_t) n);
    double* mags = (double*) malloc(sizeof(double) * (size_t) (n / 2 + 1));
    double max_mag = 0.0;
    const double* chans[2] = { ir_l, ir_r };
    for (int c = 0; c < 2; ++c)
    {
        const double* src = chans[c];
        if (!src) continue;
        for (int i = 0; i < n; ++i) { re[i] = i < copy_len ? src[i] * win[i] : 0.0; im[i] = 0.0; }
        for (int i = 1, j = 0; i < n; ++i)
        {
            int bit = n >> 1;
            while (j & bit) { j ^= bit; bit >>= 1; }
            j ^= bit;
            if (i < j) { const double t = re[i]; re[i] = re[j]; re[j] = t; }
        }
        for (int l = 2; l <= n; l <<= 1)
        {
            const double ang = -2.0 * pi / (double) l, wr = cos(ang), wi = sin(ang);
            for (int i = 0; i < n; i += l)
            {
                double tr = 1.0, ti = 0.0;
                for (int j = 0; j < l / 2; ++j)
                {
                    const int i1 = i + j, i2 = i + j + l / 2;
                    const double xr = tr * re[i2] - ti * im[i2], xi = tr * im[i2] + ti * re[i2];
                    const double ur = re[i1], ui = im[i1];
                    re[i1] = ur + xr; im[i1] = ui + xi;
                    re[i2] = ur - xr; im[i2] = ui - xi;
                    const double nr = tr * wr - ti * wi, ni = tr * wi + ti * wr;
                    tr = nr; ti = ni;
                }
            }
        }
        const int nb = n / 2;
        for (int b = 0; b <= nb; ++b)
        {
            /* the CCS packing keeps only the real parts of bins 0 and N/2 */
            mags[b] = (b == 0 || b == nb) ? fabs(re[b]) : sqrt(re[b] * re[b] + im[b] * im[b]);
            if (mags[b] > max_mag) max_mag = mags[b];
        }
        for (int b = 1; b < nb - 1; ++b)
        {
            const double ym1 = mags[b - 1], y0 = mags[b], yp1 = mags[b + 1];
            if (y0 > ym1 && y0 > yp1 && y0 > 1e-18 && ym1 > 1e-18 && yp1 > 1e-18)
            {
                const double lm = log(ym1), l0 = log(y0), lp = log(yp1);
                const double den = lm - 2.0 * l0 + lp;
                if (fabs(den) > 1e-18)
                {
                    const double delta = 0.5 * (lm - lp) / den;
                    const double v = y0 * exp(-delta * (l0 - lm));
                    if (v > max_mag) max_mag = v;
                }
            }
        }
    }
    free(win); free(re); free(im); free(mags);
    max_mag /= wmean;
    return max_mag > 1e-18 ? max_mag : 1.0;
}

/* IRConverter::computeScaleFactor, IRConverter.cpp:13-196: energy normalisation with a -6 dB margin, peak / RMS / frequency
 * response clamps, and the jump protection against the IR that is playing.  out = { scaleFactor, hasScaleFactor,
 * additionalAttenuationDb (float in the reference) }.  cur_* nullable.  PARITY: the FFT stage is pinned against the
 * reference's IRAnalyzer.cpp; the arithmetic around it restates IRConverter.cpp (which needs JUCE's file classes). */
void cpqo_ir_scale_factor(const double* ir_l, const double* ir_r, int len, const double* cur_l, const double* cur_r, int cur_len,
                          double cur_scale, double* out3)
{
    out3[0] = 1.0; out3[1] = 0.0; out3[2] = 0.0;
    if (len <= 0 || !ir_l) { out3[1] = 1.0; return; }   /* computeEnergyScale returns 1.0 for an empty buffer */
    const double* chans[2] = { ir_l, ir_r };
    const int nch = ir_r ? 2 : 1;
    double max_energy = 0.0;
    for (int c = 0; c < nch; ++c)
    {
        double e = 0.0;
        for (int i = 0; i < len; ++i) e += chans[c][i] * chans[c][i];
        if (isfinite(e) && e > 1.0e-18 && e > max_energy) max_energy = e;
    }
    double scale = 1.0;
    if (max_energy > 1.0e-18 && isfinite(max_energy)) scale = (1.0 / sqrt(max_energy)) * 0.5011872336272722;
    if (scale <= 0.0 || !isfinite(scale)) return;
    double result = scale;
    out3[1] = 1.0;
    double peak = 0.0, esum = 0.0;
    for (int c = 0; c < nch; ++c)
        for (int i = 0; i < len; ++i)
        {
            const double v = chans[c][i];
            if (fabs(v) > peak) peak = fabs(v);
            esum += v * v;
        }
    const double rms = sqrt(esum / (double) (nch * len));
    const double fgain = cpqo_ir_freq_peak_gain(ir_l, ir_r, len);
    double peak_db = 0.0, rms_db = 0.0, freq_db = 0.0;
    if (peak * scale > 0.5)
    {
        const double k = 0.5 / (peak * scale);
        result *= k; scale *= k;
        peak_db = -20.0 * log10(k);
    }
    if (rms * scale > 0.25)
    {
        const double k = 0.25 / (rms * scale);
        result *= k;
        rms_db = -20.0 * log10(k);
    }
    if (fgain > 1.41)
    {
        const double k = 1.41 / fgain;
        result *= k;
        freq_db = -20.0 * log10(k);
    }
    out3[2] = (double) (float) (peak_db + rms_db + freq_db);
    if (cur_l && cur_len > 0)
    {
        const double* cc[2] = { cur_l, cur_r };
        const int cn = cur_r ? 2 : 1;
        double cpk = 0.0, cen = 0.0, npk = 0.0, nen = 0.0;
        for (int c = 0; c < cn; ++c)
            for (int i = 0; i < cur_len; ++i)
            {
                const double v = cc[c][i] * cur_scale;
                if (fabs(v) > cpk) cpk = fabs(v);
                cen += v * v;
            }
        for (int c = 0; c < nch; ++c)
            for (int i = 0; i < len; ++i)
            {
                const double v = chans[c][i] * result;
                if (fabs(v) > npk) npk = fabs(v);
                nen += v * v;
            }
        const double crms = sqrt(cen / (double) (cn * cur_len)), nrms = sqrt(nen / (double) (nch * len));
        const int pj = cpk > 1.0e-9 && npk > cpk * 4.0 && npk > 0.5, rj = crms > 1.0e-9 && nrms > crms * 4.0 && nrms > 0.25;
        if (pj || rj)
        {
            double kp = INFINITY, kr = INFINITY;
            if (npk > 1.0e-12 && cpk > 1.0e-12) kp = (cpk * 4.0) / npk;
            if (nrms > 1.0e-12 && crms > 1.0e-12) kr = (crms * 4.0) / nrms;
            const double k = kp < kr ? kp : kr;
            if (isfinite(k) && k > 0.0 && k < 1.0) result *= k;
        }
    }
    out3[0] = result;
}

/* ------------------------------------------------------------------------------------------
 * 20-band EQ.  eqprocessor/EQProcessor.{Coefficients,Processing,ProcessingCache}.cpp
 * ------------------------------------------------------------------------------------------ */
static void cpqo_bypass_coeffs(double* c) { c[0] = 1.0; c[1] = 0.0; c[2] = 0.0; c[3] = 1.0; c[4] = 0.0; c[5] = 0.0; }

/* calcSVFCoeffs + validateAndClampParameters + 5 designers, EQProcessor.Coefficients.cpp:84-130,431-618.
 * Parameters are float and are promoted to double after clamping.
 * out = {a1,a2,a3,m0,m1,m2}; type 0 LowShelf, 1 Peaking, 2 HighShelf, 3 LowPass, 4 HighPass. */
void cpqo_eq_design(int type, float freq, float gain_db, float q, double sr, double* out)
{
    if (sr <= 0.0) { cpqo_bypass_coeffs(out); return; }
    const float nyq = (float) (sr * 0.5);
    float maxf = nyq * 0.95f;
    if (maxf > 20000.0f) maxf = 20000.0f;
    freq = freq < 20.0f ? 20.0f : (maxf < freq ? maxf : freq);
    q = q < 0.01f ? 0.01f : (20.0f < q ? 20.0f : q);
    gain_db = gain_db < -48.0f ? -48.0f : (48.0f < gain_db ? 48.0f : gain_db);
    const double f = (double) freq, gdb = (double) gain_db, Q = (double) q;
    double A = 1.0, g, k;
    const double t = tan(CPQO_PI * f / sr);
    switch (type)
    {
        case 0: A = pow(10.0, gdb / 40.0); g = t / sqrt(A); k = 1.0 / Q; break;
        case 1: A = pow(10.0, gdb / 40.0); g = t; k = 1.0 / (Q * A); break;
        case 2: A = pow(10.0, gdb / 40.0); g = t * sqrt(A); k = 1.0 / Q; break;
        case 3: g = t; k = 1.0 / Q; break;
        case 4: g = t; k = 1.0 / Q; break;
        default: memset(out, 0, sizeof(double) * 6); return;
    }
    if (!isfinite(g) || !isfinite(k)) { cpqo_bypass_coeffs(out); return; }
    const double den = 1.0 + g * (g + k);
    if (fabs(den) < 1.0e-15) { cpqo_bypass_coeffs(out); return; }
    const double a1 = 1.0 / den, a2 = g * a1, a3 = g * a2;
    out[0] = a1; out[1] = a2; out[2] = a3;
    switch (type)
    {
        case 0: out[3] = 1.0; out[4] = k * (A - 1.0); out[5] = A * A - 1.0; break;
        case 1: out[3] = 1.0; out[4] = (A - 1.0 / A) / Q; out[5] = 0.0; break;
        case 2: out[3] = A * A; out[4] = k * (1.0 - A) * A; out[5] = 1.0 - A * A; break;
        case 3: out[3] = 0.0; out[4] = 0.0; out[5] = 1.0; break;
        default: out[3] = 1.0; out[4] = -k; out[5] = -1.0; break;
    }
}

/* Decibels::decibelsToGain<double>((double)dbFloat) as used by storeTotalGainDb (EQProcessor.h:447-451) */
double cpqo_db_to_gain(float db)
{
    const double d = (double) db;
    return d > -100.0 ? pow(10.0, d * 0.05) : 0.0;
}

/* fastTanh 27/9 Pade: scalar variant returns +-1 outside +-4.5 (FastTanhApprox.h:101-107);
 * SSE variant clamps x to +-4.5 and evaluates (FastTanhApprox.h:112-119). */
static double cpqo_tanh_scalar(double x)
{
    if (x >= 4.5) return 1.0;
    if (x <= -4.5) return -1.0;
    const double x2 = x * x;
    return x * (27.0 + x2) / (27.0 + 9.0 * x2);
}
static double cpqo_tanh_sse(double x)
{
    /* _mm_max_pd(x, lo) returns lo when x is NaN; then min with hi */
    double xc = (x > -4.5) ? x : -4.5;
    xc = (xc < 4.5) ? xc : 4.5;
    const double x2 = xc * xc;
    return xc * (27.0 + x2) / (27.0 + 9.0 * x2);
}
static int cpqo_valid(double v) { return isfinite(v) && fabs(v) >= 0.0 && fabs(v) < 1.0e15; }

/* processBand (mono/scalar) EQProcessor.Processing.cpp:128-186, and processBandStereo (:191-276) per lane.
 * stereo_variant selects the SSE tanh and the FMA association of the stereo kernel. */
static void cpqo_band(double* data, long n, const double* c, double* state, double sat, int stereo_variant)
{
    double ic1 = state[0], ic2 = state[1];
    const double a1 = c[0], a2 = c[1], a3 = c[2], m0 = c[3], m1 = c[4], m2 = c[5];
    for (long i = 0; i < n; ++i)
    {
        const double v0 = data[i];
        const double v3 = v0 - ic2;
        double v1, v2, out;
        if (stereo_variant)
        {
            v1 = fma(a1, ic1, a2 * v3);
            v2 = fma(a2, ic1, fma(a3, v3, ic2));
            ic1 = fma(2.0, v1, -ic1);
            ic2 = fma(2.0, v2, -ic2);
            out = fma(m0, v0, fma(m1, v1, m2 * v2));
            if (sat > 0.0) out = out * (1.0 - sat) + cpqo_tanh_sse(out) * sat;
            if (!cpqo_valid(out)) out = 0.0;
            if (!cpqo_valid(ic1)) ic1 = 0.0;
            if (!cpqo_valid(ic2)) ic2 = 0.0;
            out = out > -100.0 ? out : -100.0; /* _mm_max_pd then _mm_min_pd */
            out = out < 100.0 ? out : 100.0;
            data[i] = out;
        }
        else
        {
            v1 = a1 * ic1 + a2 * v3;
            v2 = ic2 + a2 * ic1 + a3 * v3;
            ic1 = 2.0 * v1 - ic1;
            ic2 = 2.0 * v2 - ic2;
            out = m0 * v0 + m1 * v1 + m2 * v2;
            if (sat > 0.0) out = out * (1.0 - sat) + cpqo_tanh_scalar(out) * sat;
            if (!cpqo_valid(out)) out = 0.0;
            data[i] = out < -100.0 ? -100.0 : (out > 100.0 ? 100.0 : out);
            if (!cpqo_valid(ic1)) ic1 = 0.0;
            if (!cpqo_valid(ic2)) ic2 = 0.0;
        }
    }
    state[0] = ic1; /* block-end killDenormal is the identity in Release (DspNumericPolicy.h:189-196) */
    state[1] = ic2;
}

/* LinearRamp, DspNumericPolicy.h:319-421 */
typedef struct { double current, target, step; int remaining, total_steps; } cpqo_ramp;

typedef struct
{
    double coeffs[20][6];
    int active[20];      /* EQCoeffCache::bandActive = enabled && sr > 0 (ProcessingCache.cpp:56-96) */
    int node_active[20]; /* BandNode::active = enabled && !(0-dB skip) (Coefficients.cpp:27-58); used by the node path only */
    int mode[20];        /* 0 Stereo, 1 Left, 2 Right, 3 Mid, 4 Side */
    double saturation;   /* already promoted: (double)(float) */
    double state[4][20][2]; /* filterState[kFilterChannels]: L, R, Mid, Side (EQProcessor.h:155,637) */
    cpqo_ramp gain;
    double total_gain_target; /* linear; EQProcessor::totalGainTarget */
    int structure;       /* 0 Serial, 1 Parallel */
    int agc;             /* EQParameters::agcEnabled */
    double sr;
    double agc_env_in, agc_env_out, agc_gain; /* rtAgcEnvInputShadow / OutputShadow / CurrentGainShadow */
} cpqo_eq;

cpqo_eq* cpqo_eq_create(double sr, float total_gain_db)
{
    cpqo_eq* e = (cpqo_eq*) calloc(1, sizeof(cpqo_eq));
    const int steps = (int) (sr * 0.05 + 0.5); /* SMOOTHING_TIME_SEC = 0.05, computeTotalSteps */
    e->gain.total_steps = steps > 0 ? steps : 1;
    e->gain.current = e->gain.target = cpqo_db_to_gain(total_gain_db);
    e->total_gain_target = e->gain.current;
    e->saturation = (double) 0.2f;
    e->sr = sr;
    e->agc_gain = 1.0;
    return e;
}
void cpqo_eq_destroy(cpqo_eq* e) { free(e); }

void cpqo_eq_set_band(cpqo_eq* e, int band, const double* coeffs6, int active, int mode)
{
    memcpy(e->coeffs[band], coeffs6, sizeof(double) * 6);
    e->active[band] = active;
    e->node_active[band] = active;
    e->mode[band] = mode;
}
/* BandNode::active for the node path; differs from `active` by createBandNode's 0-dB skip (Coefficients.cpp:49-53) */
void cpqo_eq_set_node_active(cpqo_eq* e, int band, int node_active) { e->node_active[band] = node_active; }
/* EQParameters::filterStructure (0 Serial, 1 Parallel) and agcEnabled */
void cpqo_eq_set_mode(cpqo_eq* e, int structure, int agc) { e->structure = structure; e->agc = agc; }
void cpqo_eq_get_agc(const cpqo_eq* e, double* out3) { out3[0] = e->agc_env_in; out3[1] = e->agc_env_out; out3[2] = e->agc_gain; }
void cpqo_eq_get_ms_state(const cpqo_eq* e, double* out) { memcpy(out, e->state[2], sizeof(double) * 2 * 20 * 2); }
void cpqo_eq_set_saturation(cpqo_eq* e, float sat) { e->saturation = (double) sat; }
void cpqo_eq_set_total_gain(cpqo_eq* e, float db) { e->total_gain_target = cpqo_db_to_gain(db); }
void cpqo_eq_get_state(const cpqo_eq* e, double* out) { memcpy(out, e->state, sizeof(double) * 2 * 20 * 2); }

/* calculateRMS, Processing.cpp:21-52 (AVX2 build: four FMA lanes, lane sum left to right, scalar remainder) */
static double cpqo_rms(const double* d, int n)
{
    if (!d || n <= 0) return 0.0;
    double lane[4] = { 0.0, 0.0, 0.0, 0.0 };
    int i = 0;
    const int vend = n / 4 * 4;
    for (; i < vend; i += 4)
        for (int k = 0; k < 4; ++k) lane[k] = fma(d[i + k], d[i + k], lane[k]);
    double sum = lane[0] + lane[1] + lane[2] + lane[3];
    for (; i < n; ++i) sum += d[i] * d[i];
    return sqrt(sum / (double) n);
}

/* processAGC + calculateAGCGain, Processing.cpp:343-445; block-rate coefficients 1 - exp(-n / (sr * tau)) from the
 * tables prepareToPlay fills (EQProcessor.Core.cpp:776-785; AGC_ATTACK/RELEASE/SMOOTH_TIME_SEC = 0.2 / 2.0 / 0.2) */
static void cpqo_agc(cpqo_eq* e, double* L, double* R, int n, double input_rms)
{
    const double dn = (double) n;
    const double att = 1.0 - exp(-dn / (e->sr * 0.2)), rel = 1.0 - exp(-dn / (e->sr * 2.0)), smo = 1.0 - exp(-dn / (e->sr * 0.2));
    double in_rms = input_rms, out_rms = 0.0;
    const double rl = cpqo_rms(L, n);
    if (rl > out_rms) out_rms = rl;
    if (R)
    {
        const double rr = cpqo_rms(R, n);
        if (rr > out_rms) out_rms = rr;
    }
    if (!isfinite(in_rms) || in_rms > 1000.0) in_rms = 1000.0;
    if (!isfinite(out_rms) || out_rms > 1000.0) out_rms = 1000.0;
    double env_in = e->agc_env_in, env_out = e->agc_env_out, cur = e->agc_gain;
    if (!isfinite(env_in)) env_in = 0.0;
    if (!isfinite(env_out)) env_out = 0.0;
    if (!isfinite(cur)) cur = 1.0;
    const double a_in = in_rms > env_in ? att : rel, a_out = out_rms > env_out ? att : rel;
    env_in = env_in * (1.0 - a_in) + in_rms * a_in;
    env_out = env_out * (1.0 - a_out) + out_rms * a_out;
    if (env_in < 1.0e-20) env_in = 0.0;
    if (env_out < 1.0e-20) env_out = 0.0;
    double target = 1.0;
    if (!(env_out < 1e-6))
    {
        const double ratio = env_in / env_out;
        if (ratio > 1.0 / 1.059 && ratio < 1.059) target = 1.0;
        else
        {
            const double lo = (double) 0.06f, hi = (double) 16.0f;   /* jlimit(AGC_MIN_GAIN, AGC_MAX_GAIN, ratio) */
            target = ratio < lo ? lo : (hi < ratio ? hi : ratio);
        }
    }
    const double next = cur * (1.0 - smo) + target * smo;
    e->agc_env_in = env_in;
    e->agc_env_out = env_out;
    e->agc_gain = next;
    const double inc = (next - cur) / dn;
    for (int i = 0; i < n; ++i)
    {
        const double g = cur + (double) i * inc;
        L[i] *= g;
        if (R) R[i] *= g;
    }
}

/* EQProcessor::process(block, params, cache): Processing.cpp:1019-1276 (Serial :1231-1253, Parallel :1132-1228, AGC
 * :1119-1131 + processAGC, total-gain ramp :1262-1274).  When an active band is Mid/Side the reference falls back to
 * the node path process(block) (:1037-1044 -> :484-1017): BandNode::active decides which bands run, Mid/Side bands are
 * encoded / processed / decoded per band (:690-740; inside the Parallel structure :790-832, from the band input, with
 * accum += work - src).  The node path's structure cross-fade (:866-940) only runs when the structure changes while
 * playing; a prepared engine starts in its requested structure.  Returns 0. */
int cpqo_eq_process(cpqo_eq* e, double* L, double* R, long total, int block)
{
    const int nch = R ? 2 : 1;
    int node_path = 0;
    for (int b = 0; b < 20; ++b)
        if (e->active[b] && e->mode[b] >= 3) node_path = 1;

    const int* on = node_path ? e->node_active : e->active;
    double* src = (double*) malloc(sizeof(double) * (size_t) block * 6);
    double *srcL = src, *srcR = src + block, *workL = src + 2 * block, *workR = src + 3 * block, *accL = src + 4 * block,
           *accR = src + 5 * block;
    for (long pos = 0; pos < total; pos += block)
    {
        const int n = (int) ((total - pos) < block ? (total - pos) : block);
        double* bl = L + pos;
        double* br = R ? R + pos : NULL;
        double input_rms = 0.0;
        if (e->agc)
        {
            const double r0 = cpqo_rms(bl, n);
            if (r0 > input_rms) input_rms = r0;
            if (br)
            {
                const double r1 = cpqo_rms(br, n);
                if (r1 > input_rms) input_rms = r1;
            }
        }
        if (e->structure == 1)
        {
            memcpy(srcL, bl, sizeof(double) * (size_t) n);
            if (br) memcpy(srcR, br, sizeof(double) * (size_t) n);
            memset(accL, 0, sizeof(double) * (size_t) n);
            memset(accR, 0, sizeof(double) * (size_t) n);
            for (int b = 0; b < 20; ++b)
            {
                if (!on[b]) continue;
                const int mode = e->mode[b];
                if (mode == 3 || mode == 4)
                {
                    if (nch < 2)
                    {
                        if (mode == 3)   /* Mono -> Mid: the band on the one channel with the Mid state */
                        {
                            memcpy(workL, srcL, sizeof(double) * (size_t) n);
                            cpqo_band(workL, n, e->coeffs[b], e->state[2][b], e->saturation, 0);
                            for (int i = 0; i < n; ++i) { accL[i] = accL[i] + workL[i]; accL[i] = accL[i] - srcL[i]; }
                        }
                        continue;
                    }
                    double* ms = (double*) malloc(sizeof(double) * (size_t) n * 2);
                    for (int i = 0; i < n; ++i)
                    {
                        ms[i] = (srcL[i] + srcR[i]) * 0.5;
                        ms[n + i] = (srcL[i] - srcR[i]) * 0.5;
                    }
                    if (mode == 3) cpqo_band(ms, n, e->coeffs[b], e->state[2][b], e->saturation, 0);
                    else cpqo_band(ms + n, n, e->coeffs[b], e->state[3][b], e->saturation, 0);
                    for (int i = 0; i < n; ++i)
                    {
                        workL[i] = ms[i] + ms[n + i];
                        workR[i] = ms[i] - ms[n + i];
                    }
                    for (int i = 0; i < n; ++i)
                    {
                        accL[i] += workL[i] - srcL[i];
                        accR[i] += workR[i] - srcR[i];
                    }
                    free(ms);
                    continue;
                }
                const int doL = (mode == 0 || mode == 1), doR = (mode == 0 || mode == 2) && nch > 1;
                const int sv = (mode == 0 && nch >= 2);
                if (doL)
                {
                    memcpy(workL, srcL, sizeof(double) * (size_t) n);
                    cpqo_band(workL, n, e->coeffs[b], e->state[0][b], e->saturation, sv);
                    for (int i = 0; i < n; ++i) { accL[i] = accL[i] + workL[i]; accL[i] = accL[i] - srcL[i]; }
                }
                if (doR)
                {
                    memcpy(workR, srcR, sizeof(double) * (size_t) n);
                    cpqo_band(workR, n, e->coeffs[b], e->state[1][b], e->saturation, sv);
                    for (int i = 0; i < n; ++i) { accR[i] = accR[i] + workR[i]; accR[i] = accR[i] - srcR[i]; }
                }
            }
            for (int i = 0; i < n; ++i) bl[i] = srcL[i] + accL[i];
            if (br)
                for (int i = 0; i < n; ++i) br[i] = srcR[i] + accR[i];
        }
        else
        {
            for (int b = 0; b < 20; ++b)
            {
                if (!on[b]) continue;
                const int mode = e->mode[b];
                if (mode == 0 && nch >= 2)
                {
                    cpqo_band(bl, n, e->coeffs[b], e->state[0][b], e->saturation, 1);
                    cpqo_band(br, n, e->coeffs[b], e->state[1][b], e->saturation, 1);
                }
                else if (mode == 3 || mode == 4)
                {
                    if (nch < 2)
                    {
                        if (mode == 3) cpqo_band(bl, n, e->coeffs[b], e->state[2][b], e->saturation, 0);   /* Mono -> Mid */
                        else memset(bl, 0, sizeof(double) * (size_t) n);                                    /* Mono -> Side = 0 */
                        continue;
                    }
                    for (int i = 0; i < n; ++i)
                    {
                        workL[i] = (bl[i] + br[i]) * 0.5;   /* msWork[0..n) = Mid */
                        workR[i] = (bl[i] - br[i]) * 0.5;   /* msWork[n..2n) = Side */
                    }
                    if (mode == 3) cpqo_band(workL, n, e->coeffs[b], e->state[2][b], e->saturation, 0);
                    else cpqo_band(workR, n, e->coeffs[b], e->state[3][b], e->saturation, 0);
                    for (int i = 0; i < n; ++i)
                    {
                        bl[i] = workL[i] + workR[i];
                        br[i] = workL[i] - workR[i];
                    }
                }
                else
                {
                    if (mode == 0 || mode == 1) cpqo_band(bl, n, e->coeffs[b], e->state[0][b], e->saturation, 0);
                    if ((mode == 0 || mode == 2) && nch > 1) cpqo_band(br, n, e->coeffs[b], e->state[1][b], e->saturation, 0);
                }
            }
        }
        if (e->agc)
        {
            cpqo_agc(e, bl, br, n, input_rms);
            continue;
        }
        /* total gain ramp, :1262-1274 + applyGainRamp_AVX2 :279-337 (gain(i) = start + i*inc) */
        cpqo_ramp* r = &e->gain;
        if (fabs(r->target - e->total_gain_target) > 1e-6 && e->total_gain_target != r->target)
        {
            r->target = e->total_gain_target;
            const int steps = r->remaining > 0 ? r->remaining : r->total_steps;
            r->step = (r->target - r->current) / (double) steps;
            r->remaining = steps;
        }
        const double start = r->current;
        if (n > 0 && r->remaining > 0)
        {
            if (n >= r->remaining) { r->current = r->target; r->remaining = 0; }
            else { r->current += r->step * (double) n; r->remaining -= n; }
        }
        const double inc = (r->current - start) / (double) n;
        for (int ch = 0; ch < nch; ++ch)
        {
            double* d = (ch == 0 ? L : R) + pos;
            for (int i = 0; i < n; ++i) d[i] *= start + (double) i * inc;
        }
    }
    free(src);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Output epilogue.  AudioEngine.Processing.DSPCoreDouble.cpp:465-469 (makeup gain), :581,:655-663
 * (headroom, no dither), PsychoacousticDither.h:192-355 (dither + 12-tap error feedback).
 * The uniform stream is injected (two per channel-sample, u1 then u2), see SURVEY fact 8.
 * ------------------------------------------------------------------------------------------ */
static const double cpqo_ns_coeffs[6][3][12] = {
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

/* coefficient selection of PsychoacousticDither::prepare, :231-275 */
void cpqo_dither_coeffs(double sample_rate, int bit_depth, double* out12)
{
    int sr_band;
    if (sample_rate < 46050.0) sr_band = 0;
    else if (sample_rate < 72000.0) sr_band = 1;
    else if (sample_rate < 144000.0) sr_band = 2;
    else if (sample_rate < 264600.0) sr_band = 3;
    else if (sample_rate < 529200.0) sr_band = 4;
    else sr_band = 5;
    const int bp = bit_depth <= 16 ? 0 : (bit_depth <= 24 ? 1 : 2);
    memcpy(out12, cpqo_ns_coeffs[sr_band][bp], sizeof(double) * 12);
}

/* One channel of the epilogue, in place.  bit_depth <= 0: y = x*makeup*0.8912509381337456.
 * Otherwise the dither/noise-shaper recurrence of processStereoBlock (:293-405) for this channel.
 * z[12] is the error history (carried); uniforms holds 2*n values; tmp_out (nullable) receives the
 * pre-quantiser value.
 *
 * The recurrence is CHAOTIC: the error-feedback filter's gains (|c_k| up to 11.7) amplify a one-ulp difference in tmp
 * until the quantiser flips, within a few hundred samples, and from there the outputs differ by whole LSBs.  Parity with
 * the reference is therefore all or nothing, and this function reproduces the association of the reference as built by
 * oracle/Makefile (g++ -O2 -mfma, which contracts a*b + c into fused multiply-adds; read off the disassembly of
 * processStereoBlock in oracle/_ref):
 *   shaped = fma(c11,z11, ... fma(c2,z2, fma(c0,z0, c1*z1)) ...)
 *   left channel of a stereo block :  tmp = fma(x, headroom, d) + shaped,          d = ((u1-0.5)+(u2-0.5))*scale
 *   right channel / mono block     :  tmp = fma((u1-0.5)+(u2-0.5), scale, x*headroom) + shaped
 * `role`: 0 = left channel of a stereo block, 1 = right channel of a stereo block or the mono path.
 * Pinned bit-for-bit against PsychoacousticDither.h compiled in place (tests/test_oracle_vs_ref.py). */
void cpqo_epilogue_ex(double* data, long n, double makeup_gain, double sample_rate, int bit_depth,
                      const double* uniforms, double* z, double* tmp_out, int role)
{
    const double headroom = 0.8912509381337456;
    if (bit_depth <= 0)
    {
        for (long i = 0; i < n; ++i) data[i] = (data[i] * makeup_gain) * headroom;
        return;
    }
    double c[12];
    cpqo_dither_coeffs(sample_rate, bit_depth, c);
    const double scale = 1.0 / pow(2.0, bit_depth - 1);
    const double inv_scale = pow(2.0, bit_depth - 1);
    for (long i = 0; i < n; ++i)
    {
        const double x = data[i] * makeup_gain;
        double shaped = c[1] * z[1];
        shaped = fma(c[0], z[0], shaped);
        for (int t = 2; t < 12; ++t) shaped = fma(c[t], z[t], shaped);
        const double u1 = uniforms[2 * i], u2 = uniforms[2 * i + 1];
        const double tpdf = (u1 - 0.5) + (u2 - 0.5);
        const double tmp = (role == 0 ? fma(x, headroom, tpdf * scale) : fma(tpdf, scale, x * headroom)) + shaped;
        const double q = nearbyint(tmp * inv_scale) * scale; /* round-half-even, default FP env */
        double err = tmp - q;
        if (fabs(err) < 1.0e-20) err = 0.0;
        for (int t = 11; t > 0; --t) z[t] = z[t - 1];
        z[0] = err;
        if (tmp_out) tmp_out[i] = tmp;
        data[i] = q;
    }
}
/* The uniform stream of channel `channel` of PsychoacousticDither(seed) when its VSL stream is not valid: SplitMix64(seed)
 * hands channel i its seedValue (:118-140), fallbackState = seedValue ^ 0xd1b54a32d192ed03, then xorshift64* per draw
 * (fallbackUniform, :485-497).  Writes n uniforms (two per sample: u1, u2). */
void cpqo_dither_fallback_uniforms(unsigned long long seed, int channel, long n, double* out)
{
    unsigned long long sm = seed, x = 0;
    for (int i = 0; i <= channel; ++i)
    {
        unsigned long long z = (sm += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        x = (z ^ (z >> 31)) ^ 0xd1b54a32d192ed03ULL;
    }
    for (long i = 0; i < n; ++i)
    {
        x ^= x >> 12;
        x ^= x << 25;
        x ^= x >> 27;
        out[i] = (double) ((x * 2685821657736338717ULL) >> 11) * (1.0 / 9007199254740992.0);
    }
}

void cpqo_epilogue(double* data, long n, double makeup_gain, double sample_rate, int bit_depth,
                   const double* uniforms, double* z, double* tmp_out)
{
    cpqo_epilogue_ex(data, n, makeup_gain, sample_rate, bit_depth, uniforms, z, tmp_out, 1);
}

/* The ConvolverThenEQ chain per callback (DSPCoreDouble.cpp:386-414,465-469,655-663), no dither. */
void cpqo_chain_process(cpqo_nuc* nl, cpqo_nuc* nr, cpqo_eq* eq, double* L, double* R, long total, int block,
                        int outer, double makeup, int epilogue)
{
    for (long pos = 0; pos < total; pos += block)
    {
        const int n = (int) ((total - pos) < block ? (total - pos) : block);
        double* ch[2] = { L + pos, R ? R + pos : NULL };
        cpqo_nuc* cv[2] = { nl, nr };
        for (int c = 0; c < (R ? 2 : 1); ++c)
        {
            if (!cv[c]) continue;
            cpqo_nuc_process(cv[c], ch[c], ch[c], n, n);
            if (outer) cpqo_outer_wet(ch[c], n, 1.0);
        }
        if (eq) cpqo_eq_process(eq, ch[0], ch[1], n, n);
        if (epilogue)
            for (int c = 0; c < (R ? 2 : 1); ++c) cpqo_epilogue(ch[c], n, makeup, 48000.0, 0, NULL, NULL, NULL);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Output stages of DSPCore::processDouble around the path: OutputFilter (src/OutputFilter.cpp), makeup gain
 * (AudioEngine.Processing.DSPCoreDouble.cpp:465-469), output DC blocker (src/UltraHighRateDCBlocker.h), headroom,
 * scrub and hard clamp (DSPCoreDouble.cpp:655-691, 712-737).  SimplePeakLimiter (:700-710) is the identity for
 * |y| <= 0.787 and is not restated.
 * ------------------------------------------------------------------------------------------------ */
static void cpqo_biquad_identity(double* c) { c[0] = 1.0; c[1] = c[2] = c[3] = c[4] = 0.0; }

/* OutputFilter::makeLPF / makeHPF, OutputFilter.cpp:28-70: RBJ cookbook, normalised by a0; {b0,b1,b2,a1,a2} */
static void cpqo_biquad_lpf(double fc, double q, double fs, double* c)
{
    const double nyq = fs * 0.4999;
    if (fc >= nyq || q <= 0.0 || fs <= 0.0) { cpqo_biquad_identity(c); return; }
    const double w0 = 2.0 * 3.14159265358979323846 * fc / fs;
    const double sn = sin(w0), cs = cos(w0);
    const double alpha = sn / (2.0 * q);
    const double a0inv = 1.0 / (1.0 + alpha);
    c[0] = (1.0 - cs) * 0.5 * a0inv;
    c[1] = (1.0 - cs) * a0inv;
    c[2] = (1.0 - cs) * 0.5 * a0inv;
    c[3] = (-2.0 * cs) * a0inv;
    c[4] = (1.0 - alpha) * a0inv;
}
static void cpqo_biquad_hpf(double fc, double q, double fs, double* c)
{
    const double nyq = fs * 0.4999;
    if (fc <= 0.0 || fc >= nyq || q <= 0.0 || fs <= 0.0) { cpqo_biquad_identity(c); return; }
    const double w0 = 2.0 * 3.14159265358979323846 * fc / fs;
    const double sn = sin(w0), cs = cos(w0);
    const double alpha = sn / (2.0 * q);
    const double a0inv = 1.0 / (1.0 + alpha);
    c[0] = (1.0 + cs) * 0.5 * a0inv;
    c[1] = -(1.0 + cs) * a0inv;
    c[2] = (1.0 + cs) * 0.5 * a0inv;
    c[3] = (-2.0 * cs) * a0inv;
    c[4] = (1.0 - alpha) * a0inv;
}

/* The three stages OutputFilter::process cascades, in processing order (prepare :72-112, process :160-166, :290-296). */
void cpqo_out_design(double sr, int conv_is_last, int hc, int lc, int lp, double* out15)
{
    const double fc_hc = (sr <= 48000.0) ? 19000.0 : 22000.0;
    const double fc_lp = (sr <= 48000.0) ? 19000.0 : 24000.0;
    if (conv_is_last)
    {
        if (lc == 1) cpqo_biquad_hpf(15.0, 0.5, sr, out15);
        else cpqo_biquad_hpf(18.0, 0.70711, sr, out15);
        if (hc == 0) { cpqo_biquad_lpf(fc_hc, 0.54120, sr, out15 + 5); cpqo_biquad_lpf(fc_hc, 1.30656, sr, out15 + 10); }
        else if (hc == 2) { cpqo_biquad_lpf(fc_hc, 0.5, sr, out15 + 5); cpqo_biquad_identity(out15 + 10); }
        else { cpqo_biquad_lpf(fc_hc, 0.70711, sr, out15 + 5); cpqo_biquad_lpf(fc_hc, 0.70711, sr, out15 + 10); }
    }
    else
    {
        const double q = lp == 0 ? 1.0 : (lp == 2 ? 0.5 : 0.70711);
        cpqo_biquad_hpf(20.0, 0.70711, sr, out15);
        cpqo_biquad_lpf(fc_lp, q, sr, out15 + 5);
        cpqo_biquad_lpf(fc_lp, q, sr, out15 + 10);
    }
}

typedef struct cpqo_out
{
    double sr;
    double w[2][3][2];   /* [ch][stage][w1,w2] */
    double dc_alpha[2];
    double dc_state[2][2];
    double lim_release, lim_env;   /* SimplePeakLimiter: exp(-1 / (sr * release)), envelope; release 0 = limiter off */
    int lim_on;
} cpqo_out;

/* SimplePeakLimiter::prepare / reset (audioengine/SimplePeakLimiter.h:18-30); the engine prepares it with 100 ms
 * (AudioEngine.Processing.DSPCoreLifecycle.cpp:228).  release_ms <= 0 switches the stage off. */
void cpqo_out_set_limiter(cpqo_out* o, double release_ms)
{
    o->lim_on = release_ms > 0.0;
    const double sec = release_ms * 0.001;
    o->lim_release = (sec > 0.0 && o->sr > 0.0) ? exp(-1.0 / (o->sr * sec)) : 0.0;
    o->lim_env = 1.0;
}

/* SimplePeakLimiter::processBlock (:36-86) with the engine's constants (DSPCoreDouble.cpp:703-710): threshold
 * kOutputHeadroom - 0.5 dB, knee 1 dB; attack immediate, release towards 1; one envelope for both channels. */
static void cpqo_limiter_block(cpqo_out* o, double* L, double* R, int n)
{
    const double thr = 0.8413951287507587, knee = 0.108748, clip_start = thr - knee * 0.5;
    for (int i = 0; i < n; ++i)
    {
        const double al = fabs(L[i]), ar = R ? fabs(R[i]) : al;
        const double peak = al > ar ? al : ar;          /* jmax(a, b) = a < b ? b : a */
        const double sp = peak > 1.0e-12 ? peak : 1.0e-12;
        double want = 1.0;
        if (sp > clip_start)
        {
            if (sp <= thr)
            {
                const double t = (sp - clip_start) / knee;
                const double shape = t * t * (3.0 - 2.0 * t);
                want = 1.0 - (1.0 - thr / sp) * shape;
            }
            else want = thr / sp;
        }
        if (want < o->lim_env) o->lim_env = want;
        else o->lim_env = 1.0 + (o->lim_env - 1.0) * o->lim_release;
        L[i] *= o->lim_env;
        if (R) R[i] *= o->lim_env;
    }
}

cpqo_out* cpqo_out_create(double sr, double dc_cutoff)
{
    cpqo_out* o = (cpqo_out*) calloc(1, sizeof(cpqo_out));
    o->sr = sr;
    /* UltraHighRateDCBlocker::init, :60-90 */
    o->dc_alpha[0] = o->dc_alpha[1] = 1.0e-6;
    if (isfinite(sr) && sr > 0.0 && isfinite(dc_cutoff) && dc_cutoff > 0.0)
    {
        const double ratios[2] = { 1.0 - 0.1, 1.0 + 0.1 };
        for (int i = 0; i < 2; ++i)
        {
            const double omega = 2.0 * 3.14159265358979323846 * (dc_cutoff * ratios[i]) / sr;
            double a = -expm1(-omega);
            if (!isfinite(a) || a <= 0.0 || a >= 1.0) a = 1.0e-6;
            o->dc_alpha[i] = a;
        }
    }
    return o;
}
void cpqo_out_destroy(cpqo_out* o) { free(o); }

/* One DF2T step with the association of biquadStep128_FMA (OutputFilter.cpp:118-137), incl. its |w| < 1e-20 flush. */
static double cpqo_biquad_step(double x, const double* c, double* w)
{
    const double y = fma(c[0], x, w[0]);
    double n1 = fma(c[1], x, fma(-c[3], y, w[1]));
    double n2 = fma(-c[4], y, c[2] * x);
    if (fabs(n1) < 1.0e-20) n1 = 0.0;
    if (fabs(n2) < 1.0e-20) n2 = 0.0;
    w[0] = n1;
    w[1] = n2;
    return y;
}

/* Per callback: [OutputFilter] -> makeup -> [DC blocker] -> [headroom] -> [scrub] -> [peak limiter] -> [clamp]. */
void cpqo_out_process(cpqo_out* o, double* L, double* R, long total, int block, int use_filter, int conv_is_last, int hc, int lc,
                      int lp, double makeup, int use_dc, int headroom, int clamp)
{
    const double hr = 0.8912509381337456;
    double c[15];
    cpqo_out_design(o->sr, conv_is_last, hc, lc, lp, c);
    for (long pos = 0; pos < total; pos += block)
    {
        const int n = (int) ((total - pos) < block ? (total - pos) : block);
        double* ch[2] = { L + pos, R ? R + pos : NULL };
        for (int k = 0; k < (R ? 2 : 1); ++k)
        {
            double* d = ch[k];
            if (use_filter)
                for (int i = 0; i < n; ++i)
                {
                    double v = d[i];
                    for (int s = 0; s < 3; ++s) v = cpqo_biquad_step(v, c + 5 * s, o->w[k][s]);
                    d[i] = v;
                }
            for (int i = 0; i < n; ++i) d[i] *= makeup;
            if (use_dc)
            {
                /* UltraHighRateDCBlocker::process, :98-126 */
                double s0 = o->dc_state[k][0], s1 = o->dc_state[k][1];
                for (int i = 0; i < n; ++i)
                {
                    double x = d[i];
                    s0 = fma(o->dc_alpha[0], x - s0, s0);
                    x = x - s0;
                    s1 = fma(o->dc_alpha[1], x - s1, s1);
                    x = x - s1;
                    d[i] = x;
                }
                o->dc_state[k][0] = (isfinite(s0) && fabs(s0) < 1.0e15) ? s0 : 0.0;
                o->dc_state[k][1] = (isfinite(s1) && fabs(s1) < 1.0e15) ? s1 : 0.0;
            }
            if (headroom)
                for (int i = 0; i < n; ++i) d[i] *= hr;
            if (clamp)
                for (int i = 0; i < n; ++i)
                {
                    double v = d[i];
                    if (!(isfinite(v) && fabs(v) < 1.0e300)) v = 0.0;
                    d[i] = v;
                }
        }
        if (o->lim_on) cpqo_limiter_block(o, ch[0], ch[1], n);
        if (clamp)
            for (int k = 0; k < (R ? 2 : 1); ++k)
                for (int i = 0; i < n; ++i)
                {
                    const double v = ch[k][i];
                    ch[k][i] = v < -hr ? -hr : (v > hr ? hr : v);
                }
    }
}

int cpqo_abi_version(void) { return 2; }
