import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from oracle.bindings import best_checker
from tests import signals
chk = best_checker()
T, block, irlen = 8192, 512, 3000
irs = [signals.synth_ir(irlen, 40), signals.synth_ir(irlen, 41)]
x = np.stack([signals.noise(T, 700 + i) for i in range(4)])
want = np.stack([chk.nuc_run(irs[i % 2], x[i], block)[0] for i in range(4)])
want_sw = np.stack([chk.nuc_run(irs[(i + 1) % 2], x[i], block)[0] for i in range(4)])
for order in ((0, 1), (1, 0)):
    eng = ConvoPeqEngine(2, 2, 48000.0, block, T, shared_ir=True)
    for ch in order:
        eng.set_impulse(-1, ch, irs[ch])
    y = x.copy(); eng.process(y, capi.STAGE_CONV); eng.close()
    print(order, "err", np.abs(y - want).max(axis=1), "err vs swapped", np.abs(y - want_sw).max(axis=1), "|y|", np.abs(y).max(axis=1))
