"""
Runs the product's FFT kernel SOURCE (heffte_b200/csrc/fft_device.cuh, through fft_host_plan.h) on the CPU with the
thread-per-CUDA-thread emulation in tests/emul/ and compares with numpy: index arithmetic, digit reversal, twiddle
tables, r2c / c2r / r2r load-store modes and the strided / contiguous / generic kernel selection -- before any GPU time
is spent.  The emulation is test infrastructure; the product library never contains it.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import heffte_oracle as O
from tests.helpers import line_geometry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    from heffte_b200 import _lib
    out = os.path.join(EMUL_DIR, "_build", "libemul.so")
    sources = [os.path.join(EMUL_DIR, "emul_fft.cpp"), os.path.join(EMUL_DIR, "cuda_emul.h")] + \
              [os.path.join(ROOT, "heffte_b200", "csrc", f) for f in ("fft_device.cuh", "fft_dispatch.cuh", "fft_host_plan.h", "scatter_build.h",
                                                                      "pack_device.cuh", "pack_host.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in sources):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        # -Bsymbolic + hidden visibility: the emulated kernels carry the same C++ names as the device stubs inside
        # libheffte_b200.so (loaded RTLD_GLOBAL by other tests); they must bind to the copies in this library
        cmd = ["g++", "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread", "-fvisibility=hidden", "-Wl,-Bsymbolic", "-I", os.path.join(ROOT, "heffte_b200", "csrc"), "-I", EMUL_DIR,
               os.path.join(EMUL_DIR, "emul_fft.cpp"), "-o", out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    lib = ctypes.CDLL(out)
    lib.emul_fft1d.restype = ctypes.c_int
    lib.emul_fft1d.argtypes = [ctypes.POINTER(_lib.b200_fft1d_desc), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                               ctypes.POINTER(ctypes.c_int)]
    lib.emul_fft1d_reshape.restype = ctypes.c_int
    lib.emul_fft1d_reshape.argtypes = [ctypes.POINTER(_lib.b200_fft1d_desc), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
    lib.emul_scatter_copy.restype = ctypes.c_int
    lib.emul_scatter_copy.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def _run(emul, kind, prec, box, dim, direction, data, scale=1.0, out_box=None):
    from heffte_b200 import _lib
    g, ca, cb = line_geometry(box, dim)
    go = g if out_box is None else line_geometry(out_box, dim)[0]
    d = _lib.b200_fft1d_desc(prec, kind, box.size[dim], ca, cb, _lib.b200_line_geom(*g), _lib.b200_line_geom(*go))
    family = ctypes.c_int(-1)
    rdt, cdt = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    real_in = (kind == 1 and direction == 0) or kind >= 2
    real_out = (kind == 1 and direction == 1) or kind >= 2
    src = np.ascontiguousarray(data.astype(rdt if real_in else cdt))
    count_out = box.count() if (kind != 1 or direction == 1) else out_box.count()
    dst = np.zeros(count_out, dtype=rdt if real_out else cdt)
    rc = emul.emul_fft1d(ctypes.byref(d), direction, src.ctypes.data, dst.ctypes.data, scale, ctypes.byref(family))
    assert rc == 0
    return dst, family.value


FAMILY = {0: "strided", 1: "contig", 2: "generic", 3: "contig_real", 4: "strided_real"}


@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("n,order,dim,expect", [
    ((16, 3, 2), (0, 1, 2), 0, "contig"), ((32, 2, 2), (0, 1, 2), 0, "contig"), ((64, 2, 1), (0, 1, 2), 0, "contig"),
    ((128, 2, 1), (0, 1, 2), 0, "contig"), ((256, 1, 2), (0, 1, 2), 0, "contig"), ((512, 1, 1), (0, 1, 2), 0, "contig"),
    ((5, 16, 2), (0, 1, 2), 1, "strided"), ((9, 2, 32), (0, 1, 2), 2, "strided"), ((20, 64, 1), (0, 1, 2), 1, "strided"),
    ((10, 2, 128), (0, 1, 2), 2, "strided"), ((9, 256, 1), (0, 1, 2), 1, "strided"), ((8, 1, 512), (0, 1, 2), 2, "strided"),
    ((3, 1024, 1), (0, 1, 2), 1, "strided"), ((1024, 1, 1), (0, 1, 2), 0, "contig"),
    ((64, 3, 2), (1, 0, 2), 0, "strided"), ((3, 64, 2), (1, 2, 0), 1, "contig"),
    ((12, 5, 3), (0, 1, 2), 0, "generic"), ((4, 15, 3), (0, 1, 2), 1, "generic"), ((4, 3, 14), (2, 0, 1), 2, "generic"), ((8, 2, 2), (0, 1, 2), 0, "generic"),
])
def test_c2c_kernels_emulated(emul, prec, n, order, dim, expect):
    _check_c2c(emul, prec, n, order, dim, expect)


MIXED_LENGTHS = [96, 192, 384, 768, 1536, 80, 160, 320, 640, 1280, 200, 400, 500, 1000, 2000, 112, 224, 448, 896, 1792, 3584]
SMALL_MIXED_LENGTHS = [48, 100]      # c2c only: the real-data tables have no schedule for these halves


@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("length", MIXED_LENGTHS + SMALL_MIXED_LENGTHS)
def test_c2c_mixed_radix_kernels_emulated(emul, prec, length):
    """lengths with factors 3, 5 and 7 (radices 3, 5, 6, 7, 10, 12 in registers): contiguous and strided kernels, ragged tiles"""
    lines = 3 if length > 500 else 5
    _check_c2c(emul, prec, (length, lines, 1), (0, 1, 2), 0, "contig")
    _check_c2c(emul, prec, (lines, length, 1), (0, 1, 2), 1, "strided")
    if length in (96, 320, 1000):
        _check_c2c(emul, prec, (2, 2, length), (0, 1, 2), 2, "strided")


def _check_c2c(emul, prec, n, order, dim, expect):
    box = O.Box((0, 0, 0), tuple(v - 1 for v in n), order)
    rng = np.random.default_rng(n[0] * 7 + dim)
    x = rng.random(box.count()) + 1j * rng.random(box.count())
    tol = 2e-6 if prec == 0 else 1e-13
    y, fam = _run(emul, 0, prec, box, dim, 0, x, scale=0.5)
    assert FAMILY[fam] == expect
    assert O.rel_l2(y, 0.5 * O.exec1d_c2c(x, box, dim)) < tol
    z, _ = _run(emul, 0, prec, box, dim, 1, x)
    assert O.rel_l2(z, O.exec1d_c2c(x, box, dim, backward=True)) < tol


@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("n,order,dim", [((16, 3, 2), (0, 1, 2), 0), ((12, 5, 3), (0, 1, 2), 0), ((4, 10, 3), (0, 1, 2), 1), ((3, 2, 9), (0, 1, 2), 2),
                                         ((6, 7, 8), (1, 2, 0), 2), ((64, 2, 2), (0, 1, 2), 0), ((3, 32, 2), (0, 1, 2), 1),
                                         # power-of-two contiguous lines: the half-length complex engine (fft_contig_real_kernel)
                                         ((32, 3, 3), (0, 1, 2), 0), ((128, 5, 1), (0, 1, 2), 0), ((256, 1, 3), (0, 1, 2), 0), ((512, 2, 1), (0, 1, 2), 0),
                                         ((1024, 1, 2), (0, 1, 2), 0), ((2048, 1, 1), (0, 1, 2), 0), ((4096, 1, 1), (0, 1, 2), 0),
                                         ((3, 64, 2), (1, 2, 0), 1), ((2, 3, 128), (2, 0, 1), 2),
                                         # power-of-two lines with adjacent neighbours: fft_strided_real_kernel
                                         ((5, 32, 2), (0, 1, 2), 1), ((20, 3, 64), (0, 1, 2), 2), ((33, 128, 1), (0, 1, 2), 1), ((17, 2, 256), (0, 1, 2), 2),
                                         ((9, 512, 1), (0, 1, 2), 1), ((3, 1, 1024), (0, 1, 2), 2), ((2, 2048, 1), (0, 1, 2), 1), ((2, 4096, 1), (0, 1, 2), 1),
                                         ((64, 6, 2), (1, 0, 2), 0)])
def test_r2c_c2r_emulated(emul, prec, n, order, dim):
    box = O.Box((0, 0, 0), tuple(v - 1 for v in n), order)
    cbox = box.r2c(dim)
    rng = np.random.default_rng(n[0] + 3 * dim)
    x = rng.random(box.count())
    tol = 2e-6 if prec == 0 else 1e-13
    y, fam = _run(emul, 1, prec, box, dim, 0, x, out_box=cbox)
    size = n[dim]
    if size >= 32 and size & (size - 1) == 0:
        assert FAMILY[fam] == ("contig_real" if dim == order[0] else "strided_real")
    ref = O.exec1d_r2c(x, box, dim)
    assert O.rel_l2(y, ref) < tol
    z, _ = _run(emul, 1, prec, box, dim, 1, ref, out_box=cbox)
    assert O.rel_l2(z, O.exec1d_c2r(ref, box, dim)) < tol


@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("kind,name", [(2, "cos"), (3, "sin"), (4, "cos1")])
@pytest.mark.parametrize("n,order,dim", [((8, 3, 2), (0, 1, 2), 0), ((7, 3, 2), (0, 1, 2), 0), ((16, 2, 2), (0, 1, 2), 0), ((3, 9, 2), (1, 0, 2), 1), ((4, 2, 6), (2, 1, 0), 2),
                                         ((32, 3, 2), (0, 1, 2), 0), ((64, 5, 1), (0, 1, 2), 0), ((256, 3, 1), (0, 1, 2), 0), ((512, 1, 2), (0, 1, 2), 0),
                                         ((2048, 1, 1), (0, 1, 2), 0), ((2, 128, 3), (1, 0, 2), 1),
                                         ((5, 32, 2), (0, 1, 2), 1), ((18, 2, 64), (0, 1, 2), 2), ((33, 256, 1), (0, 1, 2), 1), ((3, 1, 512), (0, 1, 2), 2),
                                         ((2, 1024, 1), (0, 1, 2), 1), ((128, 5, 2), (2, 0, 1), 0)])
def test_r2r_emulated(emul, prec, kind, name, n, order, dim):
    box = O.Box((0, 0, 0), tuple(v - 1 for v in n), order)
    rng = np.random.default_rng(kind * 100 + n[0])
    x = rng.random(box.count())
    tol = 1e-5 if prec == 0 else 1e-12
    y, fam = _run(emul, kind, prec, box, dim, 0, x)
    size = n[dim]
    if kind != 4 and size >= 32 and size & (size - 1) == 0:
        assert FAMILY[fam] == ("contig_real" if dim == order[0] else "strided_real")
    assert O.rel_l2(y, O.r2r_forward(x, box, dim, name)) < tol
    z, _ = _run(emul, kind, prec, box, dim, 1, x)
    assert O.rel_l2(z, O.r2r_backward(x, box, dim, name)) < tol


@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("half", MIXED_LENGTHS)
def test_real_mixed_radix_kernels_emulated(emul, prec, half):
    """real transforms of length 2 * (a mixed c2c length): r2c / c2r / DCT / DST on the half-length mixed-radix engine"""
    n = 2 * half
    if half == 3584:
        pytest.skip("no real-data schedule for a line of 7168 points")
    if prec == 0 and half not in (96, 160, 500, 2000, 448):
        pytest.skip("single precision: a sample of the lengths")
    tol = 2e-5 if prec == 0 else 1e-12
    for shape, dim, family in (((n, 3, 1), 0, "contig_real"), ((3, n, 1), 1, "strided_real")):
        box = O.Box((0, 0, 0), tuple(v - 1 for v in shape))
        rng = np.random.default_rng(half + dim)
        x = rng.random(box.count())
        cbox = box.r2c(dim)
        y, fam = _run(emul, 1, prec, box, dim, 0, x, out_box=cbox)
        assert FAMILY[fam] == family
        ref = O.exec1d_r2c(x, box, dim)
        assert O.rel_l2(y, ref) < tol
        z, _ = _run(emul, 1, prec, box, dim, 1, ref, out_box=cbox)
        assert O.rel_l2(z, O.exec1d_c2r(ref, box, dim)) < tol
        for kind, name in ((2, "cos"), (3, "sin")):
            f, fam = _run(emul, kind, prec, box, dim, 0, x)
            assert FAMILY[fam] == family
            assert O.rel_l2(f, O.r2r_forward(x, box, dim, name)) < 4 * tol
            b, _ = _run(emul, kind, prec, box, dim, 1, x)
            assert O.rel_l2(b, O.r2r_backward(x, box, dim, name)) < 4 * tol
