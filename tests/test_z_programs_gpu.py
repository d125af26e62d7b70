"""
The user programs of tests/test_cpp_frontend.py (C++ front-end example, plain-C scenario of the reference's test/test_c.c) run
against the real library on the GPU: two thread-ranks on one device, one CUDA stream per rank, c2c + r2c + cosine plans.
"""
import os
import subprocess

import pytest

from tests.test_cpp_frontend import OUT, C_SRC, _compile

pytestmark = pytest.mark.gpu


def test_cpp_program_runs_on_the_gpu(built_library):
    exe = _compile(os.path.dirname(built_library), "heffte_b200", os.path.join(OUT, "example_b200"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "example_b200: ok" in r.stdout


def test_c_program_runs_on_the_gpu(built_library):
    exe = _compile(os.path.dirname(built_library), "heffte_b200", os.path.join(OUT, "test_c_b200"), C_SRC)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "test_c_b200: ok" in r.stdout
