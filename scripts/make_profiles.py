"""Turn gpurun_out captures into the tracked summaries under profiles/.

  python scripts/make_profiles.py TAG FULL.ncu-rep [LAUNCHES.csv]

FULL.ncu-rep   `ncu --set full --import-source on --profile-from-start off ... scripts/profile_step.py` (one step, one chunk)
LAUNCHES.csv   `ncu --metrics gpu__time_duration.sum --csv --profile-from-start off ... scripts/profile_step.py --streams 1024`
Writes profiles/TAG_ncu_kernels.txt (+ per-kernel stall/line summaries), profiles/TAG_launches.csv (+ shares) and
profiles/traffic.json (dram bytes per launch of every kernel, read by bench.py for roofline.traffic)."""
import collections, csv, io, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep = sys.argv[1], sys.argv[2]
launches = sys.argv[3] if len(sys.argv) > 3 else None
prof = os.path.join(root, "profiles")
os.makedirs(prof, exist_ok=True)
py = sys.executable

def run(*a):
    return subprocess.run(a, capture_output=True, text=True).stdout

summ = run(py, os.path.join(root, "scripts", "ncu_summary.py"), rep, "12")
open(os.path.join(prof, f"{tag}_ncu_kernels.txt"), "w").write(
    "# one step of scripts/profile_step.py (cfg4 geometry, one chunk of 118 sequences x 480256 samples), ncu --set full --clock-control none\n" + summ)
raw = run("ncu", "-i", rep, "--page", "raw", "--csv")
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(h)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
traffic = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("cpq::", "")
    rd = float(r[ix["dram__bytes_read.sum"]]) * scale[units[ix["dram__bytes_read.sum"]]]
    wr = float(r[ix["dram__bytes_write.sum"]]) * scale[units[ix["dram__bytes_write.sum"]]]
    traffic.setdefault(name, []).append(rd + wr)
out = {k: sum(v) / len(v) for k, v in traffic.items()}
# bench.py keys: kernel family name -> bytes per launch (average over the family's launches in one chunk)
fam = collections.defaultdict(list)
for k, v in traffic.items():
    fam[k.split("<")[0]].extend(v)
# FP64 pipe utilisation per kernel family (sm__inst_executed_pipe_fp64..pct_of_peak, as scripts/ncu_summary.py prints it): read by
# bench.py for roofline.fp64.pipe_utilisation_ncu; time-weighted over the family's launches
ncu = {}
for line in summ.splitlines():
    if not line.startswith("kernel="): continue
    f = dict(kv.split("=", 1) for kv in line.replace("kernel=void ", "kernel=").split("  ") if "=" in kv)
    fam_name = f["kernel"].split("<")[0].strip().replace("mac_tma_kernel", "mac_kernel").replace("fft_fwd16_kernel", "fft_fwd_kernel").replace("fft_inv16_kernel", "fft_inv_kernel")
    ms = float(f["ms"].replace("ms", "")); pct = float(f["fp64%"])
    a = ncu.setdefault(fam_name, [0.0, 0.0]); a[0] += ms * pct; a[1] += ms
ncu = {k: {"fp64_pct": v[0] / v[1], "ms_in_profile": v[1]} for k, v in ncu.items() if v[1] > 0}
ncu["_source"] = f"profiles/{tag}_ncu_kernels.txt (ncu --set full, one chunk)"
doc = {"_ncu": ncu, "_source": f"{os.path.basename(rep)} ({tag}): dram__bytes_read.sum + dram__bytes_write.sum per launch, chunk of 118 sequences x 480256 samples",
       "_channel_samples_per_launch": 118 * 480256,
       **{k: sum(v) / len(v) for k, v in fam.items()}, "_per_kernel": out}
for alias, real in (("mac_kernel", "mac_tma_kernel"), ("fft_fwd_kernel", "fft_fwd16_kernel"), ("fft_inv_kernel", "fft_inv16_kernel")):
    if alias not in doc and real in doc: doc[alias] = doc[real]   # the names bench.py's stage timers use
json.dump(doc, open(os.path.join(prof, "traffic.json"), "w"), indent=1)
for kern, mangled in (("eq_kernel", "_ZN3cpq9eq_kernelILb0ELb0ELb0ELb0EEEvNS_6EqArgsE"), ("mac_kernel", "_ZN3cpq10mac_kernelENS_7MacArgsE"),
                      ("mac_tma_kernel", "_ZN3cpq14mac_tma_kernelILi8EEEvNS_7MacArgsENS_12MacTensorMapES2_")):
    if kern not in fam: continue
    txt = run(py, os.path.join(root, "scripts", "ncu_stalls.py"), rep, kern, "0", "25")
    txt += "\n# per CUDA source line (needs the libcpq.so of the same build)\n" + run(py, os.path.join(root, "scripts", "ncu_lines.py"), rep, kern, mangled, "30")
    open(os.path.join(prof, f"{tag}_stalls_{kern}.txt"), "w").write(txt)
if launches:
    per = collections.OrderedDict(); lines = []
    for l in open(launches):
        if l.startswith('"'): lines.append(l)
    rr = list(csv.reader(io.StringIO("".join(lines))))
    hh = rr[0]; jx = {k: i for i, k in enumerate(hh)}
    tot = 0.0
    for r in rr[1:]:
        n = r[jx["Kernel Name"]].split("(")[0].replace("void ", "").replace("cpq::", "")
        t = float(r[jx["Metric Value"]]); tot += t
        c = per.setdefault(n, [0, 0.0]); c[0] += 1; c[1] += t
    with open(os.path.join(prof, f"{tag}_launches.csv"), "w") as f:
        f.write("# every launch of one full step (ncu gpu__time_duration.sum, cold-cache, serialised: compare shares)\n")
        f.write("kernel,launches,total_ns,share\n")
        for n, (c, t) in per.items():
            f.write(f"{n},{c},{t:.0f},{t / tot:.4f}\n")
print(open(os.path.join(prof, "traffic.json")).read())
