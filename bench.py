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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # one process per GPU, every rank streaming over PCIe in the e2e leg: keep this rank's pinned buffers and threads on
        # the GPU's own NUMA node
        from convopeq_b200.dist import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local)
    else:
        numa_cpus = None

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
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.process_device(io_dev.data_ptr(), T, T, capi.STAGE_ALL)
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1)

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
    ms_local = sum(step_ms) / len(step_ms)
    ms = torch.tensor([ms_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item())
    total_cs = float(n_seq) * T * world
    value = total_cs / (ms_per_step * 1e-3)

    # ---- e2e through the host entry point ----
    e2e = None
    if not args.no_e2e:
        host = torch.empty(n_seq, T, dtype=torch.float64).pin_memory()
        host.copy_(x_dev)
        torch.cuda.synchronize()

        def step_host():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.process_host_ptrs(host.data_ptr(), T, T, capi.STAGE_ALL)
            e1.record(stream)
            e1.synchronize()
            return e0.elapsed_time(e1)

        step_host()
        barrier()
        sampler.begin()
        e_ms = [step_host() for _ in range(args.steps)]
        barrier()
        sampler.end()
        em = torch.tensor([sum(e_ms) / len(e_ms)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(em, op=dist.ReduceOp.MAX)
        te = eng.timings()
        e2e = {"value": total_cs / (float(em.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(n_seq * T * 8 * world), "d2h_bytes_per_step": int(n_seq * T * 8 * world),
               "ms_per_step": float(em.item()), "h2d_ms": te.h2d_ms, "d2h_ms": te.d2h_ms,
               "note": "steps run in place on the pinned host buffer (each step's output is the next step's input)"}
        del host
        # informational: the same call with FP32 host buffers (hosts that hand the application float blocks): FP32 on the
        # wire, FP64 arithmetic.  Not the headline -- BASELINE's metric is the FP64 interface above.
        try:
            hostf = torch.empty(n_seq, T, dtype=torch.float32).pin_memory()
            hostf.copy_(x_dev)
            torch.cuda.synchronize()

            def step_host_f32():
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                eng.process_f32_host_ptrs(hostf.data_ptr(), T, T, capi.STAGE_ALL)
                e1.record(stream)
                e1.synchronize()
                return e0.elapsed_time(e1)

            step_host_f32()
            barrier()
            f_ms = [step_host_f32() for _ in range(min(args.steps, 3))]
            barrier()
            fm = torch.tensor([sum(f_ms) / len(f_ms)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(fm, op=dist.ReduceOp.MAX)
            e2e["f32_host_buffers"] = {"value": total_cs / (float(fm.item()) * 1e-3), "ms_per_step": float(fm.item()),
                                       "h2d_bytes_per_step": int(n_seq * T * 4 * world), "d2h_bytes_per_step": int(n_seq * T * 4 * world),
                                       "note": "cpq_process_f32: FP32 wire format, FP64 arithmetic; informational, not the headline"}
            del hostf
        except Exception as ex:   # never let the extra leg break the contract line
            e2e["f32_host_buffers"] = {"error": f"{type(ex).__name__}: {ex}"}
    clocks = sampler.stop()

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
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            # measured dram bytes per launch at the profiled launch size, scaled to this run's launch size
            traffic = tj[names[dom]] / tj["_channel_samples_per_launch"] * (cs_rank / chunks)
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
                             "achieved_tflops": flop_model[names[dom]] * cs_rank / (st[dom] * 1e-3) / 1e12,
                             "frac": flop_model[names[dom]] * cs_rank / (st[dom] * 1e-3) / 1e12 / dfma if dfma > 0 else None,
                             "note": "FP64 instruction model of the dominant kernel vs the measured DFMA probe; this path is bound by "
                                     "the FP64 pipe and shared-memory issue, not HBM (SURVEY 8d); the HBM fraction uses the compulsory "
                                     "16 B/channel-sample"}}
        cpu = None
        if not args.no_cpu and world == 1:
            v, info = cpu_reference_run(args.cpu_seconds, min(T, 96000 // BLOCK * BLOCK))
            cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(S, T), "streams_per_gpu": S, "samples_per_channel": T, "ir_taps": IR_LEN,
                       "block": BLOCK, "l2_policy": "inputs (7.9 GB/GPU at 1024 streams) far exceed the 126 MB L2; no flush needed",
                       "prepare_s": t_prep, "parallelism": f"stream-sharded x{world}, no collective",
                       "numa": (f"rank 0 bound to {len(numa_cpus)} CPUs of its GPU's NUMA node" if numa_cpus else "not bound")},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
