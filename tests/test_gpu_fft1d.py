"""
GPU parity of the batched 1-D executors (the kernels that replace cuFFT) against the oracle, through the C ABI
(include/heffte_b200_kernels.h).  Mirrors test/test_units_nompi.cpp:204-262, 336-367, 535-574 of the reference
(1-D executors on a box along each dimension, c2c / r2c / r2r, reordered boxes) at many more sizes.
"""
import ctypes

import numpy as np
import pytest

from oracle import heffte_oracle as O
from tests.helpers import TOL, seeded, line_geometry

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _exec(lib, prec, kind, box, dim, direction, x, out_count, out_dtype, scale=1.0, cbox=None):
    from heffte_b200._lib import b200_fft1d_desc, b200_line_geom
    g, ca, cb = line_geometry(box, dim)
    go = g if cbox is None else line_geometry(cbox, dim)[0]
    d = b200_fft1d_desc(prec, kind, box.size[dim], ca, cb, b200_line_geom(*g), b200_line_geom(*go))
    plan = ctypes.c_void_p()
    assert lib.b200_fft1d_create(ctypes.byref(d), ctypes.byref(plan)) == 0, lib.b200_last_error()
    xin = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    out = torch.zeros(out_count, dtype=getattr(torch, np.dtype(out_dtype).name), device="cuda")
    rc = lib.b200_fft1d_execute(plan, direction, ctypes.c_void_p(xin.data_ptr()), ctypes.c_void_p(out.data_ptr()), ctypes.c_double(scale), None)
    assert rc == 0, lib.b200_last_error()
    torch.cuda.synchronize()
    name = lib.b200_fft1d_kernel_name(plan).decode()
    lib.b200_fft1d_destroy(plan)
    return out.cpu().numpy(), name


C2C_CASES = [((n, 5, 3), 0) for n in (16, 32, 64, 128, 256, 512, 1024, 2048, 4096)] + \
            [((9, n, 2), 1) for n in (16, 32, 64, 128, 256, 512, 1024, 2048, 4096)] + \
            [((5, 7, n), 2) for n in (16, 64, 256, 512, 1024)] + \
            [((7, 6, 5), d) for d in range(3)] + [((2, 3, 4), d) for d in range(3)] + \
            [((43, 76, 24), d) for d in range(3)] + [((1, 1, 1021), 2), ((30, 1, 17), 0), ((100, 3, 1), 0), ((8, 8, 8), 1), ((3, 1, 2), 0)]


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("shape,dim", C2C_CASES)
def test_c2c_executor(lib, prec, shape, dim):
    box = O.Box((0, 0, 0), tuple(v - 1 for v in shape))
    ct = np.complex64 if prec == 0 else np.complex128
    x = seeded(box.count(), 11, True).astype(ct)
    y, name = _exec(lib, prec, 0, box, dim, 0, x, box.count(), ct)
    assert O.rel_l2(y, O.exec1d_c2c(x, box, dim)) <= TOL[prec], name
    yb, _ = _exec(lib, prec, 0, box, dim, 1, x, box.count(), ct, scale=0.25)
    assert O.rel_l2(yb, 0.25 * O.exec1d_c2c(x, box, dim, backward=True)) <= TOL[prec], name


@pytest.mark.parametrize("order", [(2, 0, 1), (1, 2, 0), (2, 1, 0)])
def test_c2c_executor_reordered_box(lib, order):
    # reference test_units_nompi.cpp:535-574 (one-dimension reorder logic)
    box = O.Box((0, 0, 0), (15, 11, 31), order)
    x = seeded(box.count(), 5, True)
    for dim in range(3):
        y, _ = _exec(lib, 1, 0, box, dim, 0, x, box.count(), np.complex128)
        assert O.rel_l2(y, O.exec1d_c2c(x, box, dim)) <= TOL[1]


def test_kernel_families(lib):
    box = O.Box((0, 0, 0), (511, 7, 7))
    x = seeded(box.count(), 3, True)
    _, name = _exec(lib, 1, 0, box, 0, 0, x, box.count(), np.complex128)
    assert name == "contig"
    box = O.Box((0, 0, 0), (7, 511, 7))
    x = seeded(box.count(), 3, True)
    _, name = _exec(lib, 1, 0, box, 1, 0, x, box.count(), np.complex128)
    assert name == "strided"


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("shape", [(8, 3, 2), (7, 3, 2), (4, 6, 5), (4, 5, 9), (64, 6, 2), (3, 16, 2), (512, 3, 3), (5, 2, 256), (2, 3, 4)])
def test_r2c_executor(lib, prec, shape):
    box = O.Box((0, 0, 0), tuple(v - 1 for v in shape))
    rt, ct = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    x = seeded(box.count(), 7, False).astype(rt)
    for dim in range(3):
        cbox = box.r2c(dim)
        y, _ = _exec(lib, prec, 1, box, dim, 0, x, cbox.count(), ct, cbox=cbox)
        ref = O.exec1d_r2c(x, box, dim)
        assert O.rel_l2(y, ref) <= TOL[prec]
        back, _ = _exec(lib, prec, 1, box, dim, 1, ref.astype(ct), box.count(), rt, cbox=cbox)
        assert O.rel_l2(back, O.exec1d_c2r(ref, box, dim)) <= TOL[prec]


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("kind", ["cos", "sin", "cos1"])
@pytest.mark.parametrize("shape", [(4, 3, 2), (5, 3, 2), (3, 7, 2), (2, 3, 16), (64, 2, 2), (6, 33, 2)])
def test_r2r_executor(lib, prec, kind, shape):
    # reference test_units_nompi.cpp:287-333 (DCT-II / DST-II known answers) generalised through the oracle
    box = O.Box((0, 0, 0), tuple(v - 1 for v in shape))
    rt = np.float32 if prec == 0 else np.float64
    x = seeded(box.count(), 9, False).astype(rt)
    kid = {"cos": 2, "sin": 3, "cos1": 4}[kind]
    for dim in range(3):
        f, _ = _exec(lib, prec, kid, box, dim, 0, x, box.count(), rt)
        assert O.rel_l2(f, O.r2r_forward(x, box, dim, kind)) <= 2 * TOL[prec]
        b, _ = _exec(lib, prec, kid, box, dim, 1, x, box.count(), rt)
        assert O.rel_l2(b, O.r2r_backward(x, box, dim, kind)) <= 2 * TOL[prec]


def test_known_answers_dct_dst(lib):
    # test/test_units_nompi.cpp:306-331
    box = O.Box((0, 0, 0), (3, 0, 0))
    x = np.array([1.0, 2.0, 3.0, 4.0])
    f, _ = _exec(lib, 1, 2, box, 0, 0, x, 4, np.float64)
    assert np.allclose(f, [20.0, -6.3086440598, 0.0, -0.4483415292], atol=1e-9)
    f, _ = _exec(lib, 1, 3, box, 0, 0, x, 4, np.float64)
    assert np.allclose(f, [13.0656296488, -5.6568542495, 5.4119610015, -4.0], atol=1e-9)


def test_large_batch_512_fp64(lib):
    """full-size lines of the headline problem (512-point fp64) on a slab, all three axes, forward then backward"""
    box = O.Box((0, 0, 0), (511, 511, 7))
    x = seeded(box.count(), 1, True)
    for dim in (0, 1):
        y, _ = _exec(lib, 1, 0, box, dim, 0, x, box.count(), np.complex128)
        assert O.rel_l2(y, O.exec1d_c2c(x, box, dim)) <= TOL[1]
    box = O.Box((0, 0, 0), (63, 15, 511))
    x = seeded(box.count(), 2, True)
    y, _ = _exec(lib, 1, 0, box, 2, 0, x, box.count(), np.complex128)
    assert O.rel_l2(y, O.exec1d_c2c(x, box, 2)) <= TOL[1]


# ---- power-of-two real transforms: the half-length complex engine (fft_contig_real_kernel / fft_strided_real_kernel) ----------
REAL_POW2_CASES = [((n, 3, 5), 0, "contig_real") for n in (32, 64, 128, 256, 512, 1024, 2048, 4096)] + \
                  [((9, n, 3), 1, "strided_real") for n in (32, 64, 128, 256, 512, 1024, 2048, 4096)] + \
                  [((37, 2, n), 2, "strided_real") for n in (32, 256, 512)] + [((1, 1, 64), 2, "contig_real"), ((512, 1, 1), 0, "contig_real")]


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("shape,dim,family", REAL_POW2_CASES)
def test_r2c_pow2_fast_kernels(lib, prec, shape, dim, family):
    box = O.Box((0, 0, 0), tuple(v - 1 for v in shape))
    rt, ct = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    x = seeded(box.count(), 13, False).astype(rt)
    cbox = box.r2c(dim)
    y, name = _exec(lib, prec, 1, box, dim, 0, x, cbox.count(), ct, scale=0.5, cbox=cbox)
    assert name == family
    ref = O.exec1d_r2c(x, box, dim)
    assert O.rel_l2(y, 0.5 * ref) <= TOL[prec]
    back, _ = _exec(lib, prec, 1, box, dim, 1, ref.astype(ct), box.count(), rt, cbox=cbox)
    assert O.rel_l2(back, O.exec1d_c2r(ref, box, dim)) <= TOL[prec]


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("kind", ["cos", "sin"])
@pytest.mark.parametrize("shape,dim,family", REAL_POW2_CASES)
def test_r2r_pow2_fast_kernels(lib, prec, kind, shape, dim, family):
    box = O.Box((0, 0, 0), tuple(v - 1 for v in shape))
    rt = np.float32 if prec == 0 else np.float64
    x = seeded(box.count(), 17, False).astype(rt)
    kid = {"cos": 2, "sin": 3}[kind]
    f, name = _exec(lib, prec, kid, box, dim, 0, x, box.count(), rt)
    assert name == family
    assert O.rel_l2(f, O.r2r_forward(x, box, dim, kind)) <= 2 * TOL[prec]
    b, _ = _exec(lib, prec, kid, box, dim, 1, x, box.count(), rt, scale=0.125)
    assert O.rel_l2(b, 0.125 * O.r2r_backward(x, box, dim, kind)) <= 2 * TOL[prec]


def test_r2c_unaligned_lines_take_the_generic_kernel(lib):
    """a real line that does not start on a complex boundary cannot be read as pairs: the plan falls back, results unchanged"""
    box = O.Box((0, 0, 0), (63, 2, 1))
    x = seeded(box.count() + 1, 21, False)
    from heffte_b200._lib import b200_fft1d_desc, b200_line_geom
    g, ca, cb = line_geometry(box, 0)
    cbox = box.r2c(0)
    go = line_geometry(cbox, 0)[0]
    d = b200_fft1d_desc(1, 1, 64, ca, cb, b200_line_geom(*g), b200_line_geom(*go))
    plan = ctypes.c_void_p()
    assert lib.b200_fft1d_create(ctypes.byref(d), ctypes.byref(plan)) == 0
    xin = torch.from_numpy(x).cuda()
    out = torch.zeros(cbox.count(), dtype=torch.complex128, device="cuda")
    rc = lib.b200_fft1d_execute(plan, 0, ctypes.c_void_p(xin.data_ptr() + 8), ctypes.c_void_p(out.data_ptr()), ctypes.c_double(1.0), None)
    assert rc == 0, lib.b200_last_error()
    torch.cuda.synchronize()
    lib.b200_fft1d_destroy(plan)
    assert O.rel_l2(out.cpu().numpy(), O.exec1d_r2c(x[1:], box, 0)) <= TOL[1]
