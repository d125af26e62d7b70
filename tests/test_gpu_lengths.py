"""
Transform lengths beyond the register / shared-memory kernels on the GPU: radix 7 (7 * 2^k), the composite engine (four-step over
sub-plans for long smooth lengths, Bluestein for primes), along all three axes, both precisions, against the oracle.  cuFFT plans any
length for the reference (include/heffte_backend_cuda.h:356-368); so must this backend -- B200_ERR_UNSUPPORTED is not an answer.
"""
import ctypes
import os
import time

import numpy as np
import pytest

from oracle import heffte_oracle as O
from tests.helpers import TOL, seeded
from tests.test_gpu_fft1d import _exec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("n", [112, 896, 3584, 8192, 16384, 10007, 12289, 1021, 7168, 6561, 30030])
def test_c2c_any_length_all_axes(lib, prec, n):
    ct = np.complex64 if prec == 0 else np.complex128
    for shape, dim in (((n, 3, 2), 0), ((4, n, 2), 1), ((3, 2, n), 2)):
        box = O.Box((0, 0, 0), tuple(v - 1 for v in shape))
        x = seeded(box.count(), 31, True).astype(ct)
        y, name = _exec(lib, prec, 0, box, dim, 0, x, box.count(), ct)
        assert O.rel_l2(y, O.exec1d_c2c(x, box, dim)) <= 4 * TOL[prec], (name, shape, dim)
        yb, _ = _exec(lib, prec, 0, box, dim, 1, x, box.count(), ct, scale=0.5)
        assert O.rel_l2(yb, 0.5 * O.exec1d_c2c(x, box, dim, backward=True)) <= 4 * TOL[prec], (name, shape, dim)


@pytest.mark.parametrize("n", [1021, 8200, 10007])
def test_real_transforms_of_awkward_lengths(lib, n):
    for shape, dim in (((n, 2, 2), 0), ((3, n, 1), 1)):
        box = O.Box((0, 0, 0), tuple(v - 1 for v in shape))
        cbox = box.r2c(dim)
        x = seeded(box.count(), 7, False)
        y, name = _exec(lib, 1, 1, box, dim, 0, x, cbox.count(), np.complex128, cbox=cbox)
        ref = O.exec1d_r2c(x, box, dim)
        assert O.rel_l2(y, ref) <= 4 * TOL[1], name
        z, _ = _exec(lib, 1, 1, box, dim, 1, ref, box.count(), np.float64, cbox=cbox)
        assert O.rel_l2(z, O.exec1d_c2r(ref, box, dim)) <= 4 * TOL[1], name
        for kind, kname in ((2, "cos"), (3, "sin")):
            f, name = _exec(lib, 1, kind, box, dim, 0, x, box.count(), np.float64)
            assert O.rel_l2(f, O.r2r_forward(x, box, dim, kname)) <= 16 * TOL[1], (name, kname)
            b, _ = _exec(lib, 1, kind, box, dim, 1, x, box.count(), np.float64)
            assert O.rel_l2(b, O.r2r_backward(x, box, dim, kname)) <= 16 * TOL[1], (name, kname)


def test_whole_plan_with_a_long_prime_axis(lib):
    """heffte::fft3d on a box whose axes take three different engines: Bluestein (10007), four-step (8192 would be too large here: 4100), radix 7"""
    import heffte_b200 as hf
    from tests.helpers import to_h
    n = (10007, 4, 14)
    world = O.world_box(n)
    x = seeded(world.count(), 3, True)
    fft = hf.fft3d(hf.backend.b200, to_h(world), to_h(world), hf.comm_self())
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty_like(dx)
    fft.forward(dx, dy, hf.scale.full)
    assert O.rel_l2(dy.cpu().numpy(), O.fft3d_forward(x, n, "c2c", scaling="full")) <= 4 * TOL[1]
    fft.backward(dy, dx)
    assert O.rel_l2(dx.cpu().numpy(), x) <= 4 * TOL[1]


def test_bluestein_beats_the_quadratic_path(lib):
    """prime 1021: the composite engine against the generic kernel (O(N p) per prime factor p), 4096 lines"""
    from heffte_b200._lib import b200_fft1d_desc, b200_line_geom
    n, lines = 1021, 4096
    x = torch.from_numpy(seeded(n * lines, 1, True)).cuda()
    y = torch.empty_like(x)
    times = {}
    for label, env in (("composite", None), ("generic", "1")):
        if env is None:
            os.environ.pop("HEFFTE_B200_NO_COMPOSITE", None)
        else:
            os.environ["HEFFTE_B200_NO_COMPOSITE"] = env
        d = b200_fft1d_desc(1, 0, n, lines, 1, b200_line_geom(1, n, 0), b200_line_geom(1, n, 0))
        plan = ctypes.c_void_p()
        assert lib.b200_fft1d_create(ctypes.byref(d), ctypes.byref(plan)) == 0
        for _ in range(2):
            lib.b200_fft1d_execute(plan, 0, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), ctypes.c_double(1.0), None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            lib.b200_fft1d_execute(plan, 0, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), ctypes.c_double(1.0), None)
        torch.cuda.synchronize()
        times[label] = (time.perf_counter() - t0) / 5
        times[label + "_kernel"] = lib.b200_fft1d_kernel_name(plan).decode()
        lib.b200_fft1d_destroy(plan)
    os.environ.pop("HEFFTE_B200_NO_COMPOSITE", None)
    print("prime 1021 x 4096 lines: composite %.3f ms (%s), generic %.3f ms (%s)" % (
        times["composite"] * 1e3, times["composite_kernel"], times["generic"] * 1e3, times["generic_kernel"]))
    assert "Bluestein" in times["composite_kernel"] and times["generic_kernel"] == "generic"
    assert times["composite"] * 3 < times["generic"]


def test_tma_loaded_tile_on_small_boxes(built_library):
    """fft_strided_tma_kernel (cp.async.bulk.tensor tile loads) is taken by default only for rows >= 1 MiB apart (the slow axis of
    512^3: covered by tests/test_y_fullsize_gpu.py and the bench parity); here it is forced on small boxes, see tests/gpu_tma_worker.py"""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_tma_worker.py")], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + "\n" + out.stderr[-3000:]
    assert "gpu_tma_worker: ok" in out.stdout
