"""World-size 2 and 4 on CPU (gloo): plan creation across ranks and the reshape routing of the product, see tests/gloo_worker.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nranks", [2, 4])
def test_gloo_ranks(built_library, nranks):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + nranks), os.path.join(ROOT, "tests", "gloo_worker.py")]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + "\n" + out.stderr[-4000:]
    assert "ok" in out.stdout
