"""PCIe ceiling on this box: pinned H2D alone, D2H alone, both at once (two streams), for the e2e roofline."""
import torch, time
n = 1 << 29   # 4 GiB of float64
h_in = torch.empty(n, dtype=torch.float64).pin_memory(); h_in.fill_(1.0)
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_a = torch.empty(n, dtype=torch.float64, device="cuda"); d_b = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
gb = n * 8 / 1e9
print(f"H2D alone  {gb / timed(h2d):6.1f} GB/s")
print(f"D2H alone  {gb / timed(d2h):6.1f} GB/s")
t = timed(both)
print(f"both       {gb / t:6.1f} GB/s each way ({2 * gb / t:6.1f} GB/s total), {t * 1e3:.1f} ms for {gb:.2f} GB each way")
# chunked (64 x 3.84 MB rows per copy, like cpq_process) both directions
rows = 480256
def chunked():
    per = 64 * rows
    for o in range(0, n - per + 1, per):
        with torch.cuda.stream(s1): d_a[o:o + per].copy_(h_in[o:o + per], non_blocking=True)
        with torch.cuda.stream(s2): h_out[o:o + per].copy_(d_b[o:o + per], non_blocking=True)
t = timed(chunked)
print(f"chunked both {gb / t:6.1f} GB/s each way")
