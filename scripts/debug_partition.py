import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from convopeq_b200.dist import partition_ranges
from oracle.bindings import best_checker, FilterSpec as OFilterSpec
from tests import signals
chk = best_checker()
for (sr5, T5, ir5, shared) in [(192000.0, 131072, 2097152, True), (192000.0, 131072, 2097152, False), (48000.0, 32768, 131072, True)]:
    block = 512
    ns = 2
    eng = ConvoPeqEngine(ns, 2, sr5, block, T5, conv_boundary=capi.CONV_INNER, shared_ir=shared)
    irs = [signals.synth_ir(ir5, 40), signals.synth_ir(ir5, 41)]
    spec = capi.default_filter_spec(sample_rate=sr5)
    for ch in range(2):
        if shared: eng.set_impulse(-1, ch, irs[ch], 1.0, spec)
        else:
            for s in range(ns): eng.set_impulse(s, ch, irs[ch], 1.0, spec)
    x5 = np.stack([signals.noise(T5, 700 + i) for i in range(2 * ns)])
    lay = eng.layout()
    parts = [lay.layers[i].num_parts_ir for i in range(lay.num_layers)]
    want = np.stack([chk.nuc_run(irs[i % 2], x5[i], block, spec=OFilterSpec(sample_rate=sr5))[0] for i in range(2 * ns)])
    y = x5.copy(); eng.set_partition_range(0, -1); eng.process(y, capi.STAGE_CONV)
    print("full", shared, parts, np.abs(y - want).max())
    for world in (2, 3):
        acc = np.zeros_like(x5)
        for r, (b, e) in enumerate(partition_ranges(parts, world)):
            eng.set_partition_range(b, e)
            y = x5.copy(); eng.process(y, capi.STAGE_CONV)
            acc += y
        print(" world", world, "err", np.abs(acc - want).max(), "per row", np.abs(acc - want).max(axis=1))
    eng.close()
