"""Compact per-kernel summary of an .ncu-rep (needs ncu on PATH): python scripts/ncu_summary.py rep [max_rows]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; maxrows = int(sys.argv[2]) if len(sys.argv) > 2 else 7
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "inst"), ("sm__cycles_elapsed.max", "cycles"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%")]
for r in rows[2:2 + maxrows]:
    out = []
    for c, n in cols:
        if c in idx:
            v = r[idx[c]]; u = units[idx[c]]
            if n == "kernel": v = v.split("(")[0][-28:]
            out.append(f"{n}={v}{'' if n in ('kernel',) else u if n in ('rd','wr','ms') else ''}")
    print("  ".join(out))
