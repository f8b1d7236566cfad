"""Tiny shapes of every kernel of the hot path in one process, for compute-sanitizer:

    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_step.py

Covers: both FFT families (16-point kernels at P = 512 / 4096, 8-point kernels at other sizes, the large-FFT passes at
P = 16384), the MAC with zero history, carried FDL rows (streaming) and a partition range, the EQ kernel in its plain /
output-stage / Parallel / statistics instantiations with chained and look-back links and tile-to-tile records, AGC, Mid/Side,
limiter, dither (cp.async double buffering), input stage, dry/wet mix, direct-form head, float conversion.  Sizes are small
because the sanitizer serialises warps; correctness of the numbers is the job of tests/, this only has to touch the code."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from tests import signals


def run(name, fn):
    fn()
    print("ok", name, flush=True)


def chain(block, ir_len, T, n_streams=2, **kw):
    eng = ConvoPeqEngine(n_streams, 2, 48000.0, block, T, conv_boundary=capi.CONV_OUTER, **kw)
    spec = capi.default_filter_spec()
    for s in range(n_streams):
        for ch in range(2):
            eng.set_impulse(s, ch, signals.synth_ir(ir_len, 3 + 2 * s + ch), 1.0, spec)
        eng.set_eq(s, signals.to_band(signals.band_params(9 + s)))
    eng.set_epilogue(1.1, 0)
    return eng


def main():
    x4 = lambda T: np.stack([signals.noise(T, 20 + i) for i in range(4)])

    def two_layers():
        T = 512 * 40          # L0 12x512 + L1 4096: two EQ tiles of 8192 + a partial one, chained links? no: 4 sequences -> look-back
        eng = chain(512, 40000, T)
        y = x4(T)
        eng.process(y, capi.STAGE_ALL)
        yf = x4(T).astype(np.float32)
        eng.process_f32(yf, capi.STAGE_ALL)
        eng.set_partition_range(3, 9)
        y = x4(T)
        eng.process(y, capi.STAGE_CONV)
        eng.close()

    def chained_links():
        os.environ["CPQ_EQ_LOOKBACK"] = "0"   # read once per process: set before the first EQ launch of this kind
        T = 512 * 40
        eng = chain(512, 6000, T)
        y = x4(T)
        eng.process(y, capi.STAGE_ALL)
        eng.close()

    def three_layers_large_fft():
        T = 256 * 160         # block 256: 256 / 2048 / 16384 (gfft passes)
        eng = ConvoPeqEngine(1, 2, 96000.0, 256, T)
        for ch in range(2):
            eng.set_impulse(0, ch, signals.synth_ir(120000, 5 + ch), 1.0, capi.default_filter_spec(sample_rate=96000.0))
        y = np.stack([signals.noise(T, 1), signals.noise(T, 2)])
        eng.process(y, capi.STAGE_CONV)
        eng.close()

    def streaming():
        T = 512 * 24
        eng = chain(512, 40000, T, n_streams=1)
        eng.set_streaming(True)
        x = np.stack([signals.noise(T, 1), signals.noise(T, 2)])
        for t0 in range(0, T, 512 * 5):
            part = np.ascontiguousarray(x[:, t0:t0 + 512 * 5])
            eng.process(part, capi.STAGE_ALL)
        blob = eng.export_state()
        eng.import_state(blob)
        eng.close()

    def eq_modes():
        T = 512 * 20
        eng = ConvoPeqEngine(3, 2, 48000.0, 512, T)
        eng.set_eq(0, signals.to_band(signals.band_params(1)), 0.2, 0.0, structure=1)                       # Parallel
        eng.set_eq(1, signals.to_band(signals.band_params(2)), 0.2, 0.0, agc=True)                           # AGC statistics
        eng.set_eq(2, signals.to_band(signals.band_params(3, modes=[(3 if i % 4 == 0 else 4 if i % 4 == 1 else 0) for i in range(20)])))  # Mid/Side
        eng.set_output_filter(True, conv_is_last=False)
        eng.set_output_stage(3.0, True)
        eng.set_peak_limiter(100.0)
        eng.set_epilogue(2.5, 0)
        y = np.stack([signals.noise(T, 30 + i, 0.4) for i in range(6)])
        eng.process(y, capi.STAGE_EQ | capi.STAGE_OUTPUT_FILTER | capi.STAGE_EPILOGUE | capi.STAGE_INPUT)
        eng.close()

    def dither():
        T = 512 * 6
        eng = ConvoPeqEngine(35, 2, 48000.0, 512, T)
        eng.set_epilogue(0.9, 24, np.random.default_rng(1).random((70, 2 * T)))
        y = np.stack([signals.noise(T, 40 + i, 0.3) for i in range(70)])
        eng.process(y, capi.STAGE_EPILOGUE)
        eng.close()

    def mix_and_head():
        T = 512 * 16
        eng = ConvoPeqEngine(1, 2, 48000.0, 512, T, conv_boundary=capi.CONV_OUTER)
        eng.set_direct_head(True)
        for ch in range(2):
            eng.set_impulse(0, ch, signals.synth_ir(9000, 7 + ch))
        eng.set_mix(0.6, 700)
        y = np.stack([signals.noise(T, 1), signals.noise(T, 2)])
        eng.process(y, capi.STAGE_CONV)
        eng.close()

    def odd_host_block():
        T = 480 * 20
        eng = ConvoPeqEngine(1, 2, 48000.0, 480, T)
        for ch in range(2):
            eng.set_impulse(0, ch, signals.synth_ir(9000, 7 + ch))
        eng.set_eq(0, signals.to_band(signals.band_params(4)))
        y = np.stack([signals.noise(T, 1), signals.noise(T, 2)])
        eng.process(y, capi.STAGE_CONV | capi.STAGE_EQ)
        eng.close()

    only = sys.argv[1:]
    for name, fn in (("two_layers", two_layers), ("three_layers_large_fft", three_layers_large_fft), ("streaming", streaming),
                     ("eq_modes", eq_modes), ("dither", dither), ("mix_and_head", mix_and_head), ("odd_host_block", odd_host_block),
                     ("chained_links", chained_links)):
        if not only or name in only:
            run(name, fn)


if __name__ == "__main__":
    main()
