#!/bin/bash
# usage: scripts/quick_bench.sh tag [extra bench args]  -> prints value / ms / stage times of a short device-only bench
tag=$1; shift
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-others "$@" > gpurun_out/qb_$tag.json 2>gpurun_out/qb_$tag.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/qb_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["value"]/1e9,3), "G/s", round(d["ms_per_step"],2), "ms", {k: round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/qb_$tag.err").read()[-1500:])
P
