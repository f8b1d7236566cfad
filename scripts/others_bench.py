"""The other_workloads legs of bench.py on their own (no cfg4 engine): python scripts/others_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
class _Skip:
    def __getattr__(self, k): raise RuntimeError("skipped")
res = bench.other_workloads(0, dev, _Skip(), torch.zeros(2, 2, device=dev), 2)
for k, v in res.items():
    if "error" in v: print(k, "ERROR", v["error"][:80]); continue
    print(f"{k:28s} device {v['device_ms']:8.3f} ms  e2e {v.get('e2e_ms', float('nan')):8.3f} ms  {v['value'] / 1e9:7.3f} G ch-samples/s  launches {v['launches_per_step']} {v.get('stage_ms')}")
