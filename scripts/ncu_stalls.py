"""Aggregate the per-instruction warp-stall samples of one kernel from an .ncu-rep (SASS view).
usage: python scripts/ncu_stalls.py rep kernel_regex [launch_index] [top_n]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# kernel_regex "id:N" selects the N-th profiled launch (1-based) instead of a name: template instances share a base name
sel = ["--kernel-id", ":::" + rx[3:]] if rx.startswith("id:") else ["--kernel-name", "regex:" + rx]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
# split per kernel instance
blocks, cur = [], []
for line in raw.splitlines():
    if line.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = []
    cur.append(line)
if cur: blocks.append(cur)
blk = blocks[which]
rows = list(csv.reader(io.StringIO("\n".join(blk[1:]))))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); per = []
ninst = 0; pipes = collections.Counter()
for r in rows[1:]:
    if len(r) < len(hdr): continue
    s = int(r[idx["# Samples"]] or 0)
    ex = int(r[idx["Instructions Executed"]] or 0)
    ninst += ex
    op = r[idx["Source"]].split()[0] if r[idx["Source"]] else "?"
    if op.startswith("@"): op = r[idx["Source"]].split()[1]
    pipes[op.split(".")[0]] += ex
    d = {h: int(r[idx[h]] or 0) for h in stalls}
    for h, v in d.items(): tot[h] += v
    per.append((s, ex, r[idx["Address"]], r[idx["Source"]], d))
allS = sum(p[0] for p in per)
print(f"kernel {blk[0]}  samples {allS}  warp-instructions {ninst}")
print("stall totals:", ", ".join(f"{h[6:]}={100*v/allS:.1f}%" for h, v in tot.most_common(10)))
print("opcode mix:", ", ".join(f"{k}={100*v/ninst:.1f}%" for k, v in pipes.most_common(18)))
print("top instructions by samples:")
for s, ex, addr, src, d in sorted(per, key=lambda p: -p[0])[:topn]:
    top = sorted(d.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {100*s/allS:5.2f}%  ex={ex:>10}  {src[:70]:70s} {' '.join(f'{h[6:]}:{v}' for h, v in top)}")
