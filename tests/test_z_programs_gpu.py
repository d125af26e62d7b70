"""
Late additions of the round (the file sorts last): sub-communicator plans on thread-ranks, and
the user programs of tests/test_cpp_frontend.py (C++ front-end example, plain-C scenario of the reference's test/test_c.c) run
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


@pytest.mark.parametrize("nranks", [4, 8])
def test_subcomm_plans_on_thread_ranks(lib, nranks):
    """plan_options::use_subcomm (test/test_subcomm.cpp): the intermediate stages live on the first ranks, the others hold empty boxes"""
    from tests.multi_rank_worker import configs
    from tests.test_gpu_threads import _run_group
    os.environ.pop("HEFFTE_B200_DISABLE_P2P", None)
    todo = [(c, 1) for c in configs(nranks, quick=True, subcomm=True) if c.get("subranks")]
    assert len(todo) == 3
    done, _ = _run_group(nranks, todo, expect_peer=True, label="_subcomm")
    assert done == len(todo)


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("length", [96, 192, 384, 768, 1536, 80, 160, 320, 640, 1280, 200, 400, 500, 1000, 2000, 48, 100, 112, 224, 448, 896, 1792, 3584])
def test_c2c_mixed_radix_lengths(lib, prec, length):
    """lengths with factors 3 and 5 on the register / shared-memory kernels (radices 3, 5, 6, 10, 12): contiguous and strided"""
    import numpy as np
    from oracle import heffte_oracle as O
    from tests.helpers import TOL, seeded
    from tests.test_gpu_fft1d import _exec
    ct = np.complex64 if prec == 0 else np.complex128
    for shape, dim, family in (((length, 5, 3), 0, "contig"), ((9, length, 2), 1, "strided"), ((7, 3, length), 2, "strided")):
        box = O.Box((0, 0, 0), tuple(v - 1 for v in shape))
        x = seeded(box.count(), 23, True).astype(ct)
        y, name = _exec(lib, prec, 0, box, dim, 0, x, box.count(), ct)
        assert name == family
        assert O.rel_l2(y, O.exec1d_c2c(x, box, dim)) <= TOL[prec], (shape, dim)
        yb, _ = _exec(lib, prec, 0, box, dim, 1, x, box.count(), ct, scale=0.5)
        assert O.rel_l2(yb, 0.5 * O.exec1d_c2c(x, box, dim, backward=True)) <= TOL[prec], (shape, dim)


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("half", [96, 192, 384, 768, 1536, 80, 160, 320, 640, 1280, 200, 400, 500, 1000, 2000, 112, 224, 448, 896, 1792])
def test_real_mixed_radix_lengths(lib, prec, half):
    """real transforms of length 2 * (a mixed c2c length): r2c / c2r / DCT / DST on the half-length mixed-radix engine"""
    import numpy as np
    from oracle import heffte_oracle as O
    from tests.helpers import TOL, seeded
    from tests.test_gpu_fft1d import _exec
    n = 2 * half
    rt, ct = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    for shape, dim, family in (((n, 3, 2), 0, "contig_real"), ((5, n, 2), 1, "strided_real")):
        box = O.Box((0, 0, 0), tuple(v - 1 for v in shape))
        x = seeded(box.count(), 29, False).astype(rt)
        cbox = box.r2c(dim)
        y, name = _exec(lib, prec, 1, box, dim, 0, x, cbox.count(), ct, cbox=cbox)
        assert name == family
        ref = O.exec1d_r2c(x, box, dim)
        assert O.rel_l2(y, ref) <= TOL[prec]
        back, _ = _exec(lib, prec, 1, box, dim, 1, ref.astype(ct), box.count(), rt, cbox=cbox)
        assert O.rel_l2(back, O.exec1d_c2r(ref, box, dim)) <= TOL[prec]
        for kid, kind in ((2, "cos"), (3, "sin")):
            f, name = _exec(lib, prec, kid, box, dim, 0, x, box.count(), rt)
            assert name == family
            assert O.rel_l2(f, O.r2r_forward(x, box, dim, kind)) <= 4 * TOL[prec]
            b, _ = _exec(lib, prec, kid, box, dim, 1, x, box.count(), rt)
            assert O.rel_l2(b, O.r2r_backward(x, box, dim, kind)) <= 4 * TOL[prec]
