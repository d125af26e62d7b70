"""
The template plug-in boundary: the REFERENCE's own test programs -- test/test_units_nompi.cpp,
test_fft3d_np{1,2,4,8}.cpp, test_fft3d_r2c.cpp, test_cos.cpp, test_reshape3d.cpp, test_streams.cpp, test_longlong.cpp,
test_subcomm.cpp -- compiled UNCHANGED (plus the `b200` twin of every `cufft` line, integration/build_reference_plugin.py)
against include/heffte_backend_b200.h, i.e. heffte::fft3d<backend::b200> as the reference's templates instantiate it:
reference planner, reference reshapes (pack -> MPI stand-in with thread-ranks -> unpack), b200 executors, packers, scaling and
device vectors from libheffte_b200.so.  The binaries are built where /root/reference exists and travel with the snapshot.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "integration", "_build")

# (program, thread-ranks, minimum number of passing b200 lines)
CASES = [("test_units_nompi", 1, 12), ("test_fft3d_np1", 1, 5), ("test_fft3d_np2", 2, 6), ("test_fft3d_np4", 4, 3), ("test_fft3d_np8", 8, 2),
         ("test_fft3d_r2c", 1, 2), ("test_fft3d_r2c", 2, 3), ("test_fft3d_r2c", 4, 2), ("test_fft3d_r2c", 8, 2),
         ("test_cos", 1, 6), ("test_cos", 2, 6), ("test_cos", 4, 6), ("test_reshape3d", 4, 12), ("test_reshape3d", 7, 0),
         ("test_streams", 6, 1), ("test_longlong", 4, 1), ("test_subcomm", 8, 1)]


def run_program(directory, name, ranks, timeout=900):
    path = os.path.join(directory, name)
    if not os.path.exists(path):
        pytest.skip("%s is not built (integration/build_reference_plugin.py needs the reference tree)" % path)
    env = dict(os.environ, SHIM_NP=str(ranks))
    out = subprocess.run([path], env=env, capture_output=True, text=True, timeout=timeout, cwd=directory)
    return out.returncode, out.stdout + out.stderr


def check_output(name, ranks, rc, text, minimum):
    assert rc == 0, "%s on %d ranks: exit code %d\n%s" % (name, ranks, rc, text[-3000:])
    assert not re.search(r"\bfail", text, re.IGNORECASE), text[-3000:]
    lines = [l for l in text.splitlines() if re.search(r"\bpass\s*$", l)]
    mine = [l for l in lines if re.search(r"b200|gpu", l)]
    assert len(mine) >= minimum, "%s on %d ranks: %d b200/gpu lines passed, expected >= %d\n%s" % (name, ranks, len(mine), minimum, text[-3000:])
    # the program's own summary line
    assert re.search(r"\S.*\s+pass\s*$", lines[-1]), text[-1000:]


@pytest.mark.gpu
@pytest.mark.parametrize("name,ranks,minimum", CASES)
def test_reference_program_with_the_b200_backend(lib, name, ranks, minimum):
    rc, text = run_program(BUILD, name, ranks)
    check_output(name, ranks, rc, text, minimum)


@pytest.mark.gpu
def test_reference_speed3d_with_the_b200_backend(lib):
    """the reference's benchmark driver itself, backend `b200`, 2 thread-ranks on one GPU (plug-in mode: the reference's reshapes)"""
    for program, backend in (("speed3d_c2c", "b200"), ("speed3d_r2c", "b200"), ("speed3d_r2r", "b200-cos")):
        path = os.path.join(BUILD, program)
        if not os.path.exists(path):
            pytest.skip("%s is not built" % path)
        out = subprocess.run([path, backend, "double", "64", "64", "64", "-n2"], env=dict(os.environ, SHIM_NP="2"), capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        m = re.search(r"Max error:\s+([0-9.eE+-]+)", out.stdout)
        assert m and float(m.group(1)) < 1e-11, out.stdout


# ---- CPU: the same programs linked against the emulated library (kernel source executed thread by thread) ------------------
@pytest.mark.parametrize("name,ranks,minimum", [("test_units_nompi", 1, 12), ("test_cos", 2, 6)])
def test_reference_program_on_the_emulated_library(name, ranks, minimum):
    if not os.path.exists("/root/reference/include/heffte.h"):
        pytest.skip("the reference tree is not present on this host")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import build_reference_plugin as B
    emul = os.path.join(BUILD, "emul")
    if not os.path.exists(os.path.join(emul, name)):
        B.build(only=[name], emulated=True)
    rc, text = run_program(emul, name, ranks)
    check_output(name, ranks, rc, text, minimum)
