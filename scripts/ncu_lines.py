"""Warp-stall samples per CUDA source line for one kernel: joins the SASS view of an .ncu-rep with nvdisasm line info
of the SAME build of libcpq.so.   python scripts/ncu_lines.py rep kernel_regex mangled_substr [top_n]"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, rx, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "convopeq_b200", "libcpq.so")], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, on = {}, None, False
for l in dis:
    if l.startswith(".text."):
        on = mangled in l and "$" not in l
        continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: line_of[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
lines = raw.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def num(x):
    try: return int(x)
    except: return 0
base = None; agg = collections.defaultdict(lambda: [0, 0, collections.Counter()]); tot = 0; inst = -1
want = int(os.environ.get("INSTANCE", "0"))      # which launch of the kernel in the report
for r in rows[1:]:
    if len(r) < len(hdr): continue
    try: addr = int(r[ix["Address"]], 16)
    except: continue
    if base is None: base = addr
    if addr == base: inst += 1       # the csv repeats the listing per kernel instance
    if inst != want: continue
    key = line_of.get(addr - base)
    s = num(r[ix["# Samples"]]); tot += s
    a = agg[key]; a[0] += s; a[1] += num(r[ix["Instructions Executed"]])
    for h in stalls: a[2][h[6:]] += num(r[ix[h]])
src = {}
def text(key):
    if not key: return "?"
    f, n = key
    if f not in src:
        p = [os.path.join(root, "convopeq_b200", "csrc", f), os.path.join(root, "include", f)]
        p = [q for q in p if os.path.exists(q)]
        src[f] = open(p[0]).read().splitlines() if p else []
    return src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
print(f"samples {tot}")
for key, (s, ex, c) in sorted(agg.items(), key=lambda kv: -kv[1][int(os.environ.get("BYINST","0"))])[:topn]:
    top = " ".join(f"{k}:{100*v/max(s,1):.0f}%" for k, v in c.most_common(3))
    print(f"{100*s/tot:5.2f}%  inst={ex:>10}  {str(key[1]) if key else '?':>4}  {text(key):90s} {top}")
