"""Generate tests/golden/golden.npz from the reference's OWN code (oracle/_ref/libcpq_ref.so = the reference's
translation units compiled in place).  Run where /root/reference is mounted:

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by tests/signals.py, so only outputs are stored.  These vectors pin the C
restatement (oracle/cpq_oracle.c) and the CUDA path on machines where the reference tree is absent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.bindings import Ref, Oracle, FilterSpec  # noqa: E402
from tests import signals  # noqa: E402
from tests.golden.cases import (CONV_CASES, EQ_CASES, CHAIN_CASES, OUTPUT_CASES, FULL_CHAIN_CASES, conv_inputs, eq_inputs,  # noqa: E402
                                chain_inputs, output_inputs, DITHER_CASES, dither_inputs)


def main():
    ref = Ref()
    orc = Oracle()   # only for the pointwise helpers (outer wet gain, headroom) that have no reference harness entry
    out = {}
    for name, c in CONV_CASES.items():
        ir, x = conv_inputs(c)
        spec = FilterSpec(**c["spec"]) if c["spec"] is not None else None
        y, lay = ref.nuc_run(ir, x, c["block"], scale=c.get("scale", 1.0), spec=spec, direct_head=c.get("direct_head", False))
        out["conv/" + name] = y
        out["conv_layout/" + name] = np.array([[l["part_size"], l["num_parts_ir"], l["parts_per_callback"], l["output_delay_samples"]]
                                               for l in lay["layers"]], dtype=np.int64)
        out["conv_gains/" + name] = np.array(lay["gains"])
    for name, c in EQ_CASES.items():
        bands, xl, xr = eq_inputs(c)
        l, r, st = ref.eq_run(signals.to_eqband(bands), xl, xr, c["sr"], c["block"], **c.get("kw", {}))
        out["eq/" + name] = np.stack([l, r])
        out["eq_state/" + name] = st
    for name, c in CHAIN_CASES.items():
        irs, bands, x = chain_inputs(c)
        y = ref.chain_run(irs, signals.to_eqband(bands), x, c["sr"], c["block"], FilterSpec(**c["spec"]), makeup=c["makeup"])
        out["chain/" + name] = y
    for name, c in OUTPUT_CASES.items():
        out["output/" + name] = ref.output_run(output_inputs(c), c["sr"], c["block"], **c["kw"])
        out["output_design/" + name] = ref.output_design(c["sr"], c["kw"].get("conv_is_last", False), c["kw"].get("hc", 1),
                                                         c["kw"].get("lc", 0), c["kw"].get("lp", 1))
    for name, c in FULL_CHAIN_CASES.items():
        # conv -> wet gain -> EQ (reference chain, no epilogue), then the reference output stages; the stages have no
        # feedback into each other, so running them one after the other over the whole signal equals the per-callback chain
        irs, bands, x = chain_inputs(c)
        y = ref.chain_run(irs, signals.to_eqband(bands), x, c["sr"], c["block"], FilterSpec(**c["spec"]), do_epilogue=False)
        out["full_chain/" + name] = ref.output_run(y, c["sr"], c["block"], makeup=c["makeup"], **c["out"])
    for name, c in DITHER_CASES.items():
        x, u = dither_inputs(c)
        q, z = ref.dither_run(x, u, c["sr"], c["bits"], c["block"])
        out["dither/" + name] = q
        out["dither_z/" + name] = z
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
