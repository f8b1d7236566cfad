"""Golden-vector case table shared by the generator and the tests (shortened versions of BASELINE.json's configs)."""
import numpy as np

from tests import signals

CONV_CASES = {
    # cfg1a: 65,536 taps, block 512 -> L0 12x512 + L1 15x4096 (g1 1.4375, D1 7168); reference nullptr FilterSpec
    "cfg1a_nospec": dict(ir_len=65536, block=512, T=16384, spec=None, seed=1),
    # cfg1 with the production default FilterSpec (HC/LC Natural)
    "cfg1_default_spec": dict(ir_len=65536, block=512, T=16384, spec={}, seed=2),
    # cfg3 geometry @96 kHz: L0 23x512 + L1 62x4096
    "cfg3_96k": dict(ir_len=262144, block=512, T=16384, spec=dict(sample_rate=96000.0), seed=3),
    # cfg4 geometry: 131,072 taps
    "cfg4_131k": dict(ir_len=131072, block=512, T=16384, spec={}, seed=4),
    # three layers 64/512/4096
    "three_layer_b64": dict(ir_len=65536, block=64, T=8192, spec=None, seed=5),
    # irregular plan (tail drops/starves): B=1024 default
    "irregular_b1024": dict(ir_len=65536, block=1024, T=32768, spec=None, seed=6),
    # air-absorption tail mode with tilt + IR scale
    "tailmode0_scaled": dict(ir_len=70000, block=512, T=16384, spec=dict(tail_mode=0, tail_start_seconds=0.03), seed=7, scale=0.37),
    # experimental direct-form head: first 32 taps as a direct FIR that bypasses the spectrum filter
    "direct_head_spec": dict(ir_len=65536, block=512, T=16384, spec={}, seed=9, scale=0.8, direct_head=True),
    # impulse at n = B-1 (layout / onset check)
    "impulse_at_511": dict(ir_len=65536, block=512, T=16384, spec=None, seed=8, impulse_at=511),
}

EQ_CASES = {
    "cfg2_default": dict(sr=48000.0, block=512, T=16384, bands=dict(seed=7)),
    "cfg2_sat0": dict(sr=48000.0, block=512, T=16384, bands=dict(seed=7), kw=dict(saturation=0.0)),
    "stress_q20": dict(sr=48000.0, block=512, T=16384, bands=dict(seed=7, stress=True)),
    "modes_types": dict(sr=96000.0, block=256, T=8192, bands=dict(seed=8, modes=[i % 3 for i in range(20)], types=[i % 5 for i in range(20)])),
    "gain_ramp": dict(sr=48000.0, block=512, T=16384, bands=dict(seed=7), kw=dict(total_gain_db=0.0, gain_change_db=-6.0, gain_change_at=512 * 8)),
    # Parallel structure, block-rate AGC, Mid/Side bands (node path, incl. createBandNode's 0-dB skip)
    "parallel": dict(sr=48000.0, block=512, T=16384, bands=dict(seed=7), kw=dict(structure=1)),
    "parallel_modes": dict(sr=96000.0, block=256, T=8192, bands=dict(seed=8, modes=[i % 3 for i in range(20)], types=[i % 5 for i in range(20)]),
                           kw=dict(structure=1, saturation=0.0)),
    "agc": dict(sr=48000.0, block=512, T=32768, bands=dict(seed=7), kw=dict(agc=True), amp=3.0),
    "agc_parallel_b64": dict(sr=44100.0, block=64, T=16384, bands=dict(seed=9), kw=dict(agc=True, structure=1), amp=0.02),
    "mid_side": dict(sr=48000.0, block=512, T=16384, bands=dict(seed=7, modes=[0, 3, 4, 1, 2] * 4, flat=[5, 6, 7])),
    "parallel_mid_side": dict(sr=48000.0, block=512, T=16384, bands=dict(seed=7, modes=[0, 3, 4, 1, 2] * 4), kw=dict(structure=1), amp=1.0),
    "mid_side_agc": dict(sr=48000.0, block=1024, T=16384, bands=dict(seed=10, modes=[3, 4] * 10), kw=dict(agc=True)),
}

CHAIN_CASES = {
    "cfg4_chain": dict(sr=48000.0, block=512, T=16384, ir_len=131072, spec={}, makeup=1.3, seed=11),
}


# output stages (OutputFilter -> makeup -> DC blocker -> headroom -> scrub + clamp) on their own, and behind conv -> EQ
OUTPUT_CASES = {
    "eq_last_natural": dict(sr=48000.0, block=512, T=16384, kw=dict(conv_is_last=False, lp=1, makeup=1.2, dc_cutoff=3.0), amp=3.0, seed=21),
    "conv_last_sharp_96k": dict(sr=96000.0, block=256, T=16384, kw=dict(conv_is_last=True, hc=0, lc=1, makeup=0.8, dc_cutoff=3.0), amp=1.0, seed=22),
    "peak_limiter": dict(sr=48000.0, block=512, T=16384, kw=dict(conv_is_last=False, lp=1, makeup=1.0, dc_cutoff=3.0, limiter_ms=100.0), amp=8.0, seed=24),
    "conv_last_soft_no_dc": dict(sr=48000.0, block=512, T=8192, kw=dict(conv_is_last=True, hc=2, lc=0, dc_cutoff=0.0, clamp=False), amp=1.0, seed=23),
}
FULL_CHAIN_CASES = {
    "cfg4_full_chain": dict(sr=48000.0, block=512, T=16384, ir_len=131072, spec={}, makeup=1.3, seed=12,
                            out=dict(conv_is_last=False, lp=1, dc_cutoff=3.0)),
}


def output_inputs(c):
    return np.stack([signals.noise(c["T"], 6000 + c["seed"]), signals.noise(c["T"], 6500 + c["seed"])]) * c["amp"] + 0.05


def conv_inputs(c):
    ir = signals.synth_ir(c["ir_len"], 1000 + c["seed"])
    if "impulse_at" in c:
        x = signals.impulse(c["T"], c["impulse_at"])
    else:
        x = signals.noise(c["T"], 2000 + c["seed"])
    return ir, x


def eq_inputs(c):
    bands = signals.band_params(**c["bands"])
    xl, xr = signals.log_sweep(c["T"], c["sr"])
    if "amp" in c:   # decorrelated channels at another level (AGC, Mid/Side)
        xl = xl * (2.0 * c["amp"]) + signals.noise(c["T"], 77, 0.05 * c["amp"])
        xr = xr * c["amp"]
    return bands, xl, xr


def chain_inputs(c):
    irs = (signals.synth_ir(c["ir_len"], 3000 + c["seed"]), signals.synth_ir(c["ir_len"], 3001 + c["seed"]))
    bands = signals.band_params(4000 + c["seed"])
    x = np.stack([signals.noise(c["T"], 5000 + c["seed"]), signals.noise(c["T"], 5001 + c["seed"])])
    return irs, bands, x


# Dither (PsychoacousticDither.h compiled in place, injected uniforms); the recurrence is chaotic, vectors are bit-exact pins.
DITHER_CASES = {
    "stereo_48k_24bit": dict(sr=48000.0, bits=24, nch=2, block=512, T=8192, seed=1),
    "stereo_96k_16bit_480": dict(sr=96000.0, bits=16, nch=2, block=480, T=9600, seed=2),
    "mono_44k1_32bit": dict(sr=44100.0, bits=32, nch=1, block=64, T=4096, seed=3),
}


def dither_inputs(c):
    x = np.stack([signals.noise(c["T"], 7000 + c["seed"] + i, 0.3) for i in range(c["nch"])])
    u = np.random.default_rng(7100 + c["seed"]).random((c["nch"], 2 * c["T"]))
    return x, u
