"""
Whole multi-rank plans on the CPU (no GPU): the emulated build of the product library with the ranks as host threads, see
tests/emul_worker.py.  Covers the peer-memory mode (fused FFT + reshape through scatter maps, fences, buffer alternation)
and the pack / exchange / unpack mode on 2, 3, 4, 6 and 8 ranks for c2c, r2c (three directions) and r2r plans.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nranks,mode,stride,env", [
    (2, "peer", 4, {}), (4, "peer", 5, {}), (8, "peer", 9, {}), (3, "peer", 3, {}), (6, "peer", 4, {}),
    (2, "exchange", 6, {}), (4, "exchange", 9, {}),
    # the switches of the executed plan: the reference's plan as it is (reorder, transposing fused reshapes),
    # one pinned decomposition, and the whole output landing in the arena before the copy
    (4, "peer", 11, {"HEFFTE_B200_REFERENCE_PLAN": "1"}),
    (4, "exchange", 13, {"HEFFTE_B200_REFERENCE_PLAN": "1"}),
    (8, "peer", 17, {"HEFFTE_B200_DECOMPOSITION": "pencils", "HEFFTE_B200_NO_DIRECT_OUTPUT": "1"}),
    (4, "peer", 7, {"HEFFTE_B200_NO_REGISTERED_OUTPUT": "1", "HEFFTE_B200_CHECK_REGISTERED": "1"}),
    # one rank: pairs of local transforms run slab by slab (one plane per slab here), and the plain path next to it
    (1, "exchange", 2, {"HEFFTE_B200_L2_SLAB_MB": "0.002"}),
    (1, "exchange", 3, {}),
])
def test_emulated_ranks(nranks, mode, stride, env):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "emul_worker.py"), str(nranks), mode, str(stride)]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1700, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stdout[-3000:] + "\n" + out.stderr[-3000:]
    assert " ok" in out.stdout


def test_registered_arrays_contract():
    """heffte_b200_register_buffer: stores straight into the registered arrays; a call that breaks the contract is refused on every
    rank when HEFFTE_B200_CHECK_REGISTERED=1 (tests/emul_registered_worker.py)"""
    cmd = [sys.executable, os.path.join(ROOT, "tests", "emul_registered_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + "\n" + out.stderr[-3000:]
    assert "emul_registered_worker: ok" in out.stdout
