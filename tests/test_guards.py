"""Memory-safety and race evidence without compute-sanitizer (closed on this pool, see profiles/sanitizer_r02.txt):

* guard mode (CPQ_GUARD=1): every device buffer of the library carries 256-byte canaries; after scripts/sanitize_step.py has
  driven every kernel family at small, awkward sizes in a child process, no canary may have been touched;
* determinism: the hand-rolled flag protocols of the EQ kernel (mbarrier mailboxes, self-flagging global records) and the MAC's
  mixed-proxy ring must give the same bits on every run -- a race would show up as run-to-run differences."""
import os
import subprocess
import sys

import numpy as np
import pytest

from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from tests import signals

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("lookback", ["", "0", "1"])
def test_no_kernel_writes_outside_its_buffers(lookback):
    env = dict(os.environ, CPQ_GUARD="1")
    if lookback:
        env["CPQ_EQ_LOOKBACK"] = lookback
    code = ("import sys; sys.path.insert(0, %r); sys.argv = ['sanitize_step.py']; import runpy\n"
            "runpy.run_path(%r, run_name='__main__')\n"
            "from convopeq_b200 import capi\n"
            "print('GUARDS', capi.load().cpq_debug_check_guards())\n") % (ROOT, os.path.join(ROOT, "scripts", "sanitize_step.py"))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "GUARDS 0" in out.stdout, out.stdout[-500:]


def test_guard_mode_is_active_in_the_child_process():
    """cpq_debug_check_guards() answers 0 (guarded and clean), not -1 (guards off), when CPQ_GUARD=1 is set."""
    env = dict(os.environ, CPQ_GUARD="1")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import numpy as np\n"
            "from convopeq_b200 import capi\n"
            "from convopeq_b200.engine import ConvoPeqEngine\n"
            "e = ConvoPeqEngine(1, 2, 48000.0, 512, 1024)\n"
            "y = np.zeros((2, 1024)); e.process(y, capi.STAGE_EPILOGUE)\n"     # allocates the guarded io buffer [2][1024]
            "L = capi.load(); assert L.cpq_debug_check_guards() == 0\n"
            "st = e.eq_state(0)\n"
            "print('CLEAN', L.cpq_debug_check_guards())\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "CLEAN 0" in out.stdout, out.stdout[-500:] + out.stderr[-1500:]


def test_runs_are_bit_identical():
    sr, block, T = 48000.0, 512, 512 * 48
    eng = ConvoPeqEngine(6, 2, sr, block, T, conv_boundary=capi.CONV_OUTER)
    spec = capi.default_filter_spec()
    for s in range(6):
        for ch in range(2):
            eng.set_impulse(s, ch, signals.synth_ir(70000, 11 + 2 * s + ch), 1.0, spec)
        eng.set_eq(s, signals.to_band(signals.band_params(50 + s)), 0.2, 0.0, structure=s % 2, agc=(s == 3))
    eng.set_output_filter(True)
    eng.set_output_stage(3.0, True)
    eng.set_epilogue(1.2, 0)
    x = np.stack([signals.noise(T, 200 + i, 0.2) for i in range(12)])
    outs = []
    for _ in range(6):
        y = x.copy()
        eng.process(y, capi.STAGE_FULL)
        outs.append(y)
    eng.close()
    for y in outs[1:]:
        assert np.array_equal(outs[0], y)
