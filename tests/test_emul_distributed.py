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


@pytest.mark.parametrize("nranks,mode,stride", [(2, "peer", 4), (4, "peer", 5), (8, "peer", 9), (3, "peer", 3), (6, "peer", 4),
                                                (2, "exchange", 6), (4, "exchange", 9)])
def test_emulated_ranks(nranks, mode, stride):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "emul_worker.py"), str(nranks), mode, str(stride)]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1700)
    assert out.returncode == 0, out.stdout[-3000:] + "\n" + out.stderr[-3000:]
    assert " ok" in out.stdout
