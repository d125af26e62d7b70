"""
Multi-GPU parity: the distributed transform on 2 / 4 / 8 ranks (one process per GPU) against the oracle, through the
reference-facing plan API.  Launches tests/multi_rank_worker.py under torch.distributed.run; skipped on boxes with a
single GPU.  Mirrors test/test_fft3d_np2.cpp ... np8.cpp, test_fft3d_r2c.cpp and test_cos.cpp of the reference.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(nranks, quick, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_rank_worker.py")] + (["--quick"] if quick else [])
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_distributed_parity(lib, nranks):
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    out = _launch(nranks, quick=(nranks != 2), port=29610 + nranks)
    assert out.returncode == 0, (out.stdout[-3000:] + "\n" + out.stderr[-3000:])
    assert "failures=0" in out.stdout
