"""
The paired kernel (two local transforms in one persistent launch) and the fused spectral-operator kernel (forward, product,
backward in one pass) executed on the CPU through the emulated library and compared with numpy; see tests/emul_kernels_worker.py.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pair_and_convolve_kernels_emulated():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul_kernels_worker.py")], cwd=ROOT, capture_output=True, text=True, timeout=1700)
    assert out.returncode == 0, out.stdout[-3000:] + "\n" + out.stderr[-3000:]
    assert "emul_kernels_worker: ok" in out.stdout
