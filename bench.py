#!/usr/bin/env python
"""bench.py -- channel-samples/s of the FP64 partitioned-convolver + 20-band EQ hot path on B200.

Workload (BASELINE.json configs[3], "cfg4"): a batch of independent stereo streams at 48 kHz, block 512,
131,072-tap IRs (distinct per stream-channel, default FilterSpec), per-stream random 20-band EQ
(saturation 0.2), makeup gain 1.0, output headroom, no dither; 10 s of noise per channel rounded up to whole
512-sample callbacks (480,256 samples).  One process per GPU; streams are sharded across ranks with no
data-path collective (weak scaling: --streams is per GPU).

  value      channel-samples/s, inputs resident in HBM (cpq_process_device), CUDA events on the engine stream
  e2e        the same through the C-ABI host entry point cpq_process: pinned host buffers, H2D + D2H inside
  roofline   dominant kernel vs MEASURED_PEAKS.json HBM GB/s using the algorithmic 16 B / channel-sample
  cpu_baseline  the reference's own CPU path (oracle/_ref) on this box's host cores, bounded sample
  other_workloads  BASELINE configs 1, 2, 3, 5 and cfg4 with 24-bit dither: device-resident and host-buffer times (rank 0)

`--impl reference` times the reference CPU path alone (rank 0 only) on the same config/metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 48000.0
BLOCK = 512
IR_LEN = 131072
T_FULL = (480000 + BLOCK - 1) // BLOCK * BLOCK     # 480256
METRIC = "channel-samples/sec FFT-conv+20-band EQ, FP64"
UNIT = "channel-samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="stereo streams per GPU")
    ap.add_argument("--samples", type=int, default=T_FULL)
    ap.add_argument("--workspace-mb", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads legs (cfg1a, cfg2, cfg3, cfg5, dither)")
    return ap.parse_args()


def band_sets(n_streams: int, seed: int):
    """Per-stream random EQ: band 0 LowShelf, 1..18 Peaking, 19 HighShelf at DEFAULT_FREQS, gains U(-6,6) dB, Q U(0.5,4)."""
    from tests import signals
    return [signals.band_params(seed + s) for s in range(n_streams)]


def workload_name(streams: int, T: int) -> str:
    return (f"cfg4: {streams} independent stereo streams/GPU x {T} samples @48kHz, block 512, 131072-tap IRs "
            f"(L0 12x512 + L1 31x4096, default FilterSpec), 20-band EQ sat 0.2, conv->EQ->makeup+headroom")


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []      # (arrival time, csv line)
        self.windows = []    # [t0, t1] timed regions (perf_counter)

    def begin(self):
        self.windows.append([time.perf_counter(), None])

    def end(self):
        self.windows[-1][1] = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        # nvidia-smi is started before the warm-up (it needs ~1 s to deliver its first line); only the samples that
        # arrived inside a timed region (device-resident steps, e2e steps) count
        for ts, ln in self.lines:
            if not any(w[0] <= ts <= (w[1] if w[1] is not None else ts) for w in self.windows):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "window": "samples taken every 100 ms inside the timed regions (device-resident steps + e2e steps)"}


# ------------------------------------------------------------------------------------------------
# reference CPU path (the checker libraries are only ever used here as the baseline being timed)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(seconds: float, T: int, max_threads: int | None = None):
    """Time the reference's per-callback ConvolverThenEQ chain on all host cores.

    One independent stream (2 convolvers + 1 EQ) per worker thread, each streaming its T-sample stereo buffer
    through the chain repeatedly (state carried, i.e. a longer signal) until `seconds` have elapsed.  ctypes
    releases the GIL inside the C call.  Returns (channel_samples_per_s, info)."""
    from oracle.bindings import best_checker, FilterSpec as OFilterSpec
    from tests import signals
    chk = best_checker()
    cores = max_threads or os.cpu_count() or 1
    spec = OFilterSpec()
    handles, bufs = [], []
    for i in range(cores):
        irs = (signals.synth_ir(IR_LEN, 5000 + 2 * i), signals.synth_ir(IR_LEN, 5001 + 2 * i))
        bands = signals.to_eqband(signals.band_params(7000 + i))
        handles.append(chk.chain_prepare(irs, bands, SR, BLOCK, spec))
        bufs.append(np.stack([signals.noise(T, 9000 + 2 * i), signals.noise(T, 9001 + 2 * i)]))
    counts = [0] * cores
    stop_at = [0.0]
    start_evt = threading.Event()

    def worker(i):
        start_evt.wait()
        while time.perf_counter() < stop_at[0]:
            chk.chain_process_prepared(handles[i], bufs[i], BLOCK, 1, 1.0, 1)
            counts[i] += 2 * T

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(cores)]
    for t in threads:
        t.start()
    t0 = time.perf_counter()
    stop_at[0] = t0 + seconds
    start_evt.set()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    for h in handles:
        chk.chain_free(h)
    total = sum(counts)
    return total / dt, {"kind": chk.kind, "cores": cores, "seconds": dt, "channel_samples": total,
                        "sample": f"{cores} independent stereo streams (one per host thread), {T} samples each, looped for "
                                  f"{dt:.1f} s = {total} channel-samples; same IR length/EQ/block as the GPU workload"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T = min(args.samples, 96000 // BLOCK * BLOCK)
    # each step = a bounded sample: all host threads stream for ~4 s
    vals, secs = [], []
    for _ in range(args.warmup):
        cpu_reference_run(0.5, T)
    info = None
    for _ in range(args.steps):
        v, info = cpu_reference_run(4.0, T)
        vals.append(v)
        secs.append(info["seconds"])
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.streams, args.samples), "reference_sample": info["sample"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# small helpers shared by the headline and the other workloads
# ------------------------------------------------------------------------------------------------
def _timed(stream, fn):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1)


def measure_workload(eng, x_dev, T, stages, steps=3, warm=2, host=True):
    """Device-resident and host-buffer time of one engine on the input x_dev [n_seq, T] (restored before every step)."""
    import torch
    dev = x_dev.device
    stream = torch.cuda.ExternalStream(eng.cuda_stream(), device=dev)
    io = torch.empty_like(x_dev)
    out = {}

    def dstep():
        io.copy_(x_dev)
        torch.cuda.synchronize()
        return _timed(stream, lambda: eng.process_device(io.data_ptr(), T, T, stages))

    for _ in range(warm):
        dstep()
    ms = [dstep() for _ in range(steps)]
    out["device_ms"] = sum(ms) / len(ms)
    cs = float(x_dev.shape[0]) * T
    out["channel_samples"] = cs
    out["value"] = cs / (out["device_ms"] * 1e-3)
    tm = eng.timings()
    out["launches_per_step"] = int(tm.kernel_launches)
    out["stage_ms"] = {"fft_fwd": round(tm.fft_fwd_ms, 4), "mac": round(tm.mac_ms, 4), "fft_inv": round(tm.fft_inv_ms, 4), "eq": round(tm.eq_ms, 4)}
    if host:
        h = torch.empty(x_dev.shape, dtype=torch.float64).pin_memory()

        def hstep():
            h.copy_(x_dev)
            torch.cuda.synchronize()
            return _timed(stream, lambda: eng.process_host_ptrs(h.data_ptr(), T, T, stages))

        hstep()
        ms = [hstep() for _ in range(steps)]
        out["e2e_ms"] = sum(ms) / len(ms)
        out["e2e_value"] = cs / (out["e2e_ms"] * 1e-3)
        del h
    del io
    return out


def other_workloads(local, dev, cfg4_eng, cfg4_x, T4):
    """BASELINE configs 1, 2, 3, 5 on one GPU, plus cfg4 with the 24-bit dither branch.  Parity of each at this size is the
    job of tests/test_full_size.py; these are the timings that go with them."""
    import torch
    from convopeq_b200 import capi
    from convopeq_b200.engine import ConvoPeqEngine
    from tests import signals
    res = {}
    g = torch.Generator(device=dev)
    g.manual_seed(77)

    def noise(n_seq, T):
        return torch.randn(n_seq, T, device=dev, dtype=torch.float64, generator=g) * 0.1

    def guard(name, fn):
        try:
            res[name] = fn()
        except Exception as ex:   # one failing leg must not take the contract line with it
            res[name] = {"error": f"{type(ex).__name__}: {ex}"}

    def cfg1(uniform):
        T = T_FULL
        eng = ConvoPeqEngine(1, 2, 48000.0, 512, T, device=local, uniform_partitions=uniform)
        for ch in range(2):
            eng.set_impulse(0, ch, signals.synth_ir(65536, 2 + ch), 1.0, None)
        r = measure_workload(eng, noise(2, T), T, capi.STAGE_CONV)
        lay = eng.layout()
        r["plan"] = " + ".join(f"{lay.layers[i].num_parts_ir}x{lay.layers[i].part_size}" for i in range(lay.num_layers))
        r["workload"] = ("cfg1b: uniform-partition extension (not a reference mode), " if uniform else "cfg1a: the reference's own plan, ") + \
            "stereo 48 kHz, 65,536-tap IR, block 512, 10 s noise, convolver only, filterSpec = nullptr"
        eng.close()
        return r

    def cfg2():
        T = 2880000 // 512 * 512
        eng = ConvoPeqEngine(1, 2, 48000.0, 512, T, device=local)
        eng.set_eq(0, signals.to_band(signals.band_params(seed=7)), 0.2, 0.0)
        xl, xr = signals.log_sweep(T, 48000.0)
        x = torch.from_numpy(np.stack([xl, xr])).to(dev)
        r = measure_workload(eng, x, T, capi.STAGE_EQ)
        r["workload"] = "cfg2: stereo 48 kHz, 20-band cascade only, 60 s log sweep (2 sequences: all parallelism from the tile chain)"
        eng.close()
        return r

    def cfg3(block):
        sr, T = 96000.0, 960000 // block * block
        eng = ConvoPeqEngine(1, 2, sr, block, T, device=local, conv_boundary=capi.CONV_OUTER)
        spec = capi.default_filter_spec(sample_rate=sr)
        for ch in range(2):
            eng.set_impulse(0, ch, signals.synth_ir(262144, 20 + ch), 1.0, spec)
        eng.set_eq(0, signals.to_band(signals.band_params(seed=7)))
        eng.set_epilogue(1.0, 0)
        r = measure_workload(eng, noise(2, T), T, capi.STAGE_ALL)
        lay = eng.layout()
        r["plan"] = " + ".join(f"{lay.layers[i].num_parts_ir}x{lay.layers[i].part_size}" for i in range(lay.num_layers))
        r["workload"] = f"cfg3: stereo 96 kHz, 262,144-tap IR, block {block}, conv -> EQ -> makeup+headroom, 10 s noise"
        eng.close()
        return r

    def cfg5():
        sr, T = 192000.0, 1920000 // 512 * 512
        eng = ConvoPeqEngine(4, 2, sr, 512, T, device=local, conv_boundary=capi.CONV_OUTER, shared_ir=True)
        spec = capi.default_filter_spec(sample_rate=sr)
        for ch in range(2):
            eng.set_impulse(-1, ch, signals.synth_ir(2097152, 40 + ch), 1.0, spec)
        for s_ in range(4):
            eng.set_eq(s_, signals.to_band(signals.band_params(seed=7 + s_)))
        eng.set_epilogue(1.0, 0)
        r = measure_workload(eng, noise(8, T), T, capi.STAGE_ALL)
        lay = eng.layout()
        r["plan"] = " + ".join(f"{lay.layers[i].num_parts_ir}x{lay.layers[i].part_size}" for i in range(lay.num_layers))
        r["workload"] = "cfg5 on ONE GPU: 192 kHz, 8 channels, 2,097,152-tap IR, block 512, conv -> EQ -> makeup+headroom, 10 s noise"
        eng.close()
        return r

    def dither():
        n_seq = cfg4_x.shape[0]
        u = torch.rand(n_seq, 2 * T4, device=dev, dtype=torch.float64, generator=g)
        cfg4_eng.set_epilogue(1.0, 24)
        cfg4_eng.set_dither_uniforms_device(u.data_ptr(), T4)
        try:
            r = measure_workload(cfg4_eng, cfg4_x, T4, capi.STAGE_ALL, steps=2, warm=1, host=False)
        finally:
            cfg4_eng.set_epilogue(1.0, 0)
        # the stage on its own (every sequence at once, 64 warps): what the serial recurrence costs whatever runs beside it
        try:
            cfg4_eng.set_epilogue(1.0, 24)
            cfg4_eng.set_dither_uniforms_device(u.data_ptr(), T4)
            r["dither_stage_alone_ms"] = measure_workload(cfg4_eng, cfg4_x, T4, capi.STAGE_EPILOGUE, steps=2, warm=1, host=False)["device_ms"]
        finally:
            cfg4_eng.set_epilogue(1.0, 0)
        r["workload"] = ("cfg4 with the 24-bit dither branch (PsychoacousticDither, injected uniforms resident on the device).  The shaper is "
                         "one dependent chain of 18 FP64-pipe operations per sample and sequence (11 DFMA, DADD, DMUL, FRND, DMUL, DADD, DSETP, "
                         "FSEL), serial in time by construction and chaotic, so it cannot be scanned or reassociated; dither_stage_alone_ms is "
                         "that chain over T samples with every sequence running at once (64 shaper warps), and the floor of the stage whatever "
                         "the batch size.  The call runs in up to four time segments through the streaming continuation, the shaper of one "
                         "segment beside the transforms of the next (CPQ_DITHER_SEGMENTS=1: one piece, the chain trails the call)")
        del u
        return r

    def streaming():
        """The streaming continuation (cpq_set_streaming) on the cfg4 batch: time per call for calls of 1 / 8 / 64 callbacks."""
        out = {}
        cfg4_eng.set_streaming(True)
        try:
            stream = torch.cuda.ExternalStream(cfg4_eng.cuda_stream(), device=dev)
            for ncb in (1, 8, 64):
                T = ncb * BLOCK
                cfg4_eng.reset()
                xs = cfg4_x[:, :T].contiguous()
                ms = []
                for _ in range(12 if ncb < 64 else 6):
                    io = xs.clone()
                    torch.cuda.synchronize()
                    ms.append(_timed(stream, lambda: cfg4_eng.process_device(io.data_ptr(), T, T, capi.STAGE_ALL)))
                med = statistics.median(ms[2:])
                out[f"{ncb}_callbacks_per_call"] = {"ms_per_call": med, "value": cfg4_x.shape[0] * T / (med * 1e-3),
                                                    "fraction_of_real_time": med * 1e-3 / (T / SR)}
        finally:
            cfg4_eng.set_streaming(False)
        out["workload"] = ("cfg4's batch (all sequences) processed in consecutive calls of 1 / 8 / 64 host callbacks with the state carried "
                           "between calls (FDL, input history, delay lines, EQ states), device-resident")
        return out

    guard("cfg1a", lambda: cfg1(False))
    guard("cfg1b_uniform_extension", lambda: cfg1(True))
    guard("cfg2", cfg2)
    guard("cfg3_block512", lambda: cfg3(512))
    guard("cfg3_block256", lambda: cfg3(256))
    guard("cfg5_one_gpu", cfg5)
    guard("cfg4_dither24", dither)
    guard("cfg4_streaming", streaming)
    return res


def pcie_duplex_ceiling(host, n_seq, T, dev, barrier, world):
    """What this box's host <-> device path delivers with nothing else going on: the whole pinned buffer down and up at the
    same time on two streams in 32 pieces (the copy pattern of cpq_process), every rank at once.  GB/s each way, whole job."""
    import torch
    import torch.distributed as dist
    d_in = torch.empty(n_seq, T, device=dev, dtype=torch.float64)
    d_out = torch.zeros(n_seq, T, device=dev, dtype=torch.float64)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    per = (n_seq + 31) // 32
    best = None
    for _ in range(3):
        barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for r0 in range(0, n_seq, per):
            with torch.cuda.stream(s1):
                d_in[r0:r0 + per].copy_(host[r0:r0 + per], non_blocking=True)
            with torch.cuda.stream(s2):
                host[r0:r0 + per].copy_(d_out[r0:r0 + per], non_blocking=True)
        e1.record(s1)
        e2.record(s2)
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), e0.elapsed_time(e2))
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t.item()) if best is None else min(best, float(t.item()))
    del d_in, d_out
    return n_seq * T * 8.0 * world / (best * 1e-3) / 1e9, best


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from convopeq_b200 import capi
    from convopeq_b200.engine import ConvoPeqEngine
    from tests import signals

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (convopeq_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # one process per GPU, every rank streaming over PCIe in the e2e leg: keep this rank's pinned buffers and threads on
        # the GPU's own NUMA node; when every GPU reports the same CPU list (single-node boxes) give each rank its own slice
        from convopeq_b200.dist import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local, local_rank=local, local_world=world)

    S, T = args.streams, args.samples // BLOCK * BLOCK
    n_seq = 2 * S
    eng = ConvoPeqEngine(n_streams=S, n_channels=2, sample_rate=SR, block_size=BLOCK, max_samples=T, device=local,
                         conv_boundary=capi.CONV_OUTER, workspace_bytes=args.workspace_mb << 20)
    # ---- synthetic IRs (device RNG -> host -> SetImpulse), per-stream EQ ----
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    decay = torch.exp(-torch.arange(IR_LEN, device=dev, dtype=torch.float64) / (IR_LEN / 6.0)) / (IR_LEN ** 0.5)
    spec = capi.default_filter_spec()
    t_prep = time.perf_counter()
    CH = 64
    for s0 in range(0, n_seq, CH):
        n = min(CH, n_seq - s0)
        irs = (torch.randn(n, IR_LEN, device=dev, dtype=torch.float64, generator=g) * decay).cpu().numpy()
        for i in range(n):
            q = s0 + i
            eng.set_impulse(q // 2, q % 2, irs[i], 1.0, spec)
    for s, params in enumerate(band_sets(S, 100000 * (rank + 1))):
        eng.set_eq(s, signals.to_band(params), 0.2, 0.0)
    eng.set_epilogue(1.0, 0)
    t_prep = time.perf_counter() - t_prep

    x_dev = torch.randn(n_seq, T, device=dev, dtype=torch.float64, generator=g) * 0.1
    io_dev = torch.empty_like(x_dev)
    stream = torch.cuda.ExternalStream(eng.cuda_stream(), device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        io_dev.copy_(x_dev)                      # restore the in-place buffer (not timed)
        torch.cuda.synchronize()
        return _timed(stream, lambda: eng.process_device(io_dev.data_ptr(), T, T, capi.STAGE_ALL))

    def allmax(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler.begin()
    ncu_range = bool(os.environ.get("CPQ_NCU_RANGE"))   # `ncu --profile-from-start off`: profile exactly the timed steps
    if ncu_range:
        torch.cuda.profiler.start()
    launches0 = eng.kernel_launch_count()
    step_ms, stage = [], []
    for _ in range(args.steps):
        step_ms.append(step_device())
        t = eng.timings()
        stage.append((t.fft_fwd_ms, t.mac_ms, t.fft_inv_ms, t.eq_ms, t.chunks))
    launches = eng.kernel_launch_count() - launches0
    barrier()
    if ncu_range:
        torch.cuda.profiler.stop()
    sampler.end()
    ms_per_step = allmax(sum(step_ms) / len(step_ms))
    total_cs = float(n_seq) * T * world
    value = total_cs / (ms_per_step * 1e-3)

    # ---- strong scaling of the same job (BASELINE cfg4 words it as 1,024 streams in total): each rank takes streams/world ----
    strong = None
    if world > 1 and S % world == 0:
        eng.set_stream_window(0, S // world)
        step_device()
        barrier()
        s_ms = [step_device() for _ in range(min(args.steps, 3))]
        barrier()
        eng.set_stream_window(0, -1)
        sm = allmax(sum(s_ms) / len(s_ms))
        strong = {"value": float(n_seq) * T / (sm * 1e-3), "unit": UNIT, "ms_per_step": sm, "streams_total": S,
                  "streams_per_gpu": S // world, "note": "strong scaling: the 1,024-stream job divided over the ranks, device-resident"}

    # ---- e2e through the host entry point ----
    e2e = None
    if not args.no_e2e:
        host = torch.empty(n_seq, T, dtype=torch.float64).pin_memory()

        def step_host():
            host.copy_(x_dev)                    # every step sees the same input as the device-resident steps (not timed)
            torch.cuda.synchronize()
            return _timed(stream, lambda: eng.process_host_ptrs(host.data_ptr(), T, T, capi.STAGE_ALL))

        step_host()
        barrier()
        sampler.begin()
        e_ms = [step_host() for _ in range(args.steps)]
        barrier()
        sampler.end()
        em = allmax(sum(e_ms) / len(e_ms))
        te = eng.timings()
        e2e = {"value": total_cs / (em * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(n_seq * T * 8 * world), "d2h_bytes_per_step": int(n_seq * T * 8 * world),
               "ms_per_step": em, "h2d_ms": te.h2d_ms, "d2h_ms": te.d2h_ms,
               "note": "cpq_process in place on a pinned host buffer that is refilled with the same input before every step"}
        # what the box's host <-> device path can do at all, every rank at once: the e2e leg is bound by it, not by a kernel
        try:
            ceil_gbs, ceil_ms = pcie_duplex_ceiling(host, n_seq, T, dev, barrier, world)
            e2e["pcie_ceiling_gbs"] = ceil_gbs
            e2e["pcie_ceiling_ms"] = ceil_ms
            e2e["achieved_gbs_each_way"] = n_seq * T * 8.0 * world / (em * 1e-3) / 1e9
            e2e["frac_of_ceiling"] = e2e["achieved_gbs_each_way"] / ceil_gbs
            e2e["pcie_note"] = ("ceiling = the same pinned buffer copied down and up simultaneously on two streams in 32 pieces by plain "
                                "cudaMemcpyAsync, all ranks at once, best of 3, whole-job GB/s each way")
        except Exception as ex:
            e2e["pcie_ceiling_error"] = f"{type(ex).__name__}: {ex}"
        del host
        # informational: a pageable host buffer (what a reference-side caller that did not pin its memory passes)
        if world == 1:
            try:
                pg = np.empty((n_seq, T), dtype=np.float64)
                pg[...] = 0.05
                first_ms = _timed(stream, lambda: eng.process(pg, capi.STAGE_ALL))   # allocates the pinned staging rings
                pg[...] = 0.05
                p_ms = _timed(stream, lambda: eng.process(pg, capi.STAGE_ALL))
                e2e["pageable_host_buffers"] = {"value": total_cs / (p_ms * 1e-3), "ms_per_step": p_ms, "first_call_ms": first_ms,
                                                "note": "cpq_process on an unpinned numpy buffer: host threads copy the rows of each sequence "
                                                        "chunk into / out of pinned staging slots beside the DMA (CPQ_STAGE_THREADS=0: the driver stages)"}
                del pg
            except Exception as ex:
                e2e["pageable_host_buffers"] = {"error": f"{type(ex).__name__}: {ex}"}
        # informational: the same call with FP32 host buffers (hosts that hand the application float blocks): FP32 on the
        # wire, FP64 arithmetic.  Not the headline -- BASELINE's metric is the FP64 interface above.
        try:
            hostf = torch.empty(n_seq, T, dtype=torch.float32).pin_memory()

            def step_host_f32():
                hostf.copy_(x_dev)
                torch.cuda.synchronize()
                return _timed(stream, lambda: eng.process_f32_host_ptrs(hostf.data_ptr(), T, T, capi.STAGE_ALL))

            step_host_f32()
            barrier()
            f_ms = [step_host_f32() for _ in range(min(args.steps, 3))]
            barrier()
            fm = allmax(sum(f_ms) / len(f_ms))
            e2e["f32_host_buffers"] = {"value": total_cs / (fm * 1e-3), "ms_per_step": fm,
                                       "h2d_bytes_per_step": int(n_seq * T * 4 * world), "d2h_bytes_per_step": int(n_seq * T * 4 * world),
                                       "note": "cpq_process_f32: FP32 wire format, FP64 arithmetic; informational, not the headline"}
            del hostf
        except Exception as ex:   # never let the extra leg break the contract line
            e2e["f32_host_buffers"] = {"error": f"{type(ex).__name__}: {ex}"}
    clocks = sampler.stop()

    others = None
    if not args.no_others and world == 1:
        others = other_workloads(local, dev, eng, x_dev, T)
    cfg5_sharded = None
    if not args.no_others and world > 1:
        try:
            from convopeq_b200.dist import bench_cfg5_sharded
            cfg5_sharded = bench_cfg5_sharded(local, rank, world)
        except Exception as ex:
            cfg5_sharded = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        st = np.mean(np.array(stage, dtype=np.float64), axis=0)
        names = ["fft_fwd_kernel", "mac_kernel", "fft_inv_kernel", "eq_kernel"]
        dom = int(np.argmax(st[:4]))
        chunks = max(int(st[4]), 1)
        launches_per_chunk = 1 if dom == 3 else 2   # two layers -> two launches of the FFT / MAC kernels per sequence chunk
        cs_rank = float(n_seq) * T
        alg_bytes_per_launch = 16.0 * cs_rank / chunks
        dur_ms = st[dom] / chunks
        achieved = alg_bytes_per_launch / (dur_ms * 1e-3) / 1e9
        traffic, ncu = None, {}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            # measured dram bytes per launch at the profiled launch size, scaled to this run's launch size
            traffic = tj[names[dom]] / tj["_channel_samples_per_launch"] * (cs_rank / chunks)
            ncu = tj.get("_ncu", {})
        except Exception:
            pass
        dfma = capi.load().cpq_probe_dfma_tflops(local, 20000)
        # FP64 instruction model per channel-sample (FMA = 2 flop): EQ 20 bands x ~14 instr, MAC 4 DFMA x (12+31)
        # taps, FFT ~ 2.5 N log2 N flop per real transform of N = 2P points -> 5 log2(2P) flop per output sample
        flop_model = {"eq_kernel": 20 * 14.0 * 2, "mac_kernel": 6.75 * (12 + 31), "fft_fwd_kernel": 5.0 * (10 + 13), "fft_inv_kernel": 5.0 * (10 + 13)}
        roofline = {"kernel": names[dom], "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes_per_launch, "launch_ms": dur_ms,
                    "launches_per_step": chunks * launches_per_chunk,
                    "stage_ms_per_step": dict(zip(names, [float(v) for v in st[:4]])),
                    "whole_path": {"achieved": 16.0 * value / world / 1e9, "frac": 16.0 * value / world / 1e9 / hbm_peak,
                                   "note": "16 B per channel-sample x channel-samples/s of the whole step on one GPU"},
                    "fp64": {"probe_dfma_tflops": dfma,
                             "pipe_utilisation_ncu": ncu.get(names[dom], {}).get("fp64_pct"),
                             "pipe_utilisation_source": ncu.get("_source"),
                             "model_achieved_tflops": flop_model[names[dom]] * cs_rank / (st[dom] * 1e-3) / 1e12,
                             "model_frac": flop_model[names[dom]] * cs_rank / (st[dom] * 1e-3) / 1e12 / dfma if dfma > 0 else None,
                             "note": "pipe_utilisation_ncu = sm__inst_executed_pipe_fp64 (ncu --set full of this kernel, committed under "
                                     "profiles/); model_* = FP64 instruction model of the dominant kernel vs the measured DFMA probe; this "
                                     "path is bound by the FP64 pipe, not HBM (SURVEY 8d); the HBM fraction uses the compulsory 16 B/channel-sample"}}
        cpu = None
        if not args.no_cpu and world == 1:
            v, info = cpu_reference_run(args.cpu_seconds, min(T, 96000 // BLOCK * BLOCK))
            cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}
            try:
                v1, info1 = cpu_reference_run(min(args.cpu_seconds, 4.0), min(T, 96000 // BLOCK * BLOCK), max_threads=1)
                cpu["one_thread"] = {"value": v1, "cores": 1, "note": "the same chain on a single host thread (BASELINE.md section 3)"}
            except Exception:
                pass
        cfg = {"workload": workload_name(S, T), "streams_per_gpu": S, "samples_per_channel": T, "ir_taps": IR_LEN,
               "block": BLOCK, "l2_policy": "inputs (7.9 GB/GPU at 1024 streams) far exceed the 126 MB L2; no flush needed",
               "prepare_s": t_prep, "parallelism": f"stream-sharded x{world}, no collective",
               "numa": (f"rank 0 bound to CPUs {numa_cpus[0]}..{numa_cpus[-1]} ({len(numa_cpus)})" if numa_cpus else "not bound")}
        if strong:
            cfg["strong_scaling"] = strong
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if others is not None:
            line["other_workloads"] = others
        if cfg5_sharded is not None:
            line["other_workloads"] = {"cfg5_partition_range_sharded": cfg5_sharded}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
