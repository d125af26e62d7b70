"""
The fused reshape (scatter maps) on the CPU: the product's host builder (csrc/scatter_build.h) and the product's kernel
source (scatter variants of the FFT kernels, scatter_copy_kernel) run under the thread emulation for EVERY rank of a
reshape; the boxes of the destination ranks are plain numpy arrays standing for local / peer GPU memory.  The result must
equal "transform, then reference reshape": oracle 1-D transform on the world followed by get_subbox per destination rank
(semantics of src/heffte_reshape3d.cpp:365-443 + include/heffte_pack3d.h:89-197).
"""
import ctypes

import numpy as np
import pytest

from oracle import heffte_oracle as O
from tests.helpers import bricks
from tests.test_emul import emul  # noqa: F401  (fixture)


def unlumped_geometry(box, dim):
    """(stride, stride_a, stride_b), count_a, count_b with a = faster of the two other axes, b = the slower one"""
    pos = box.order.index(dim)
    strides = (1, box.osize(0), box.osize(0) * box.osize(1))
    a_pos = 1 if pos == 0 else 0
    b_pos = 1 if pos == 2 else 2
    return (strides[pos], strides[a_pos], strides[b_pos]), box.osize(a_pos), box.osize(b_pos)


def _nine(boxes):
    return np.ascontiguousarray(np.array([b.nine() for b in boxes], dtype=np.int32).reshape(-1))


def _bases(arrays):
    return (ctypes.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])


CASES = [
    # world, source grid, source order, destination grid, destination order, transform dim
    ((16, 8, 6), (1, 2, 2), (0, 1, 2), (2, 1, 2), (0, 1, 2), 0),     # contig kernel, pencils dim0 -> pencils dim1
    ((6, 16, 4), (2, 1, 2), (0, 1, 2), (2, 2, 1), (0, 1, 2), 1),     # strided kernel, middle axis
    ((6, 4, 32), (2, 2, 1), (0, 1, 2), (1, 2, 2), (0, 1, 2), 2),     # strided kernel, slow axis -> bricks
    ((12, 5, 6), (1, 1, 3), (0, 1, 2), (3, 1, 1), (0, 1, 2), 0),     # generic kernel (n = 12), uneven cells
    ((16, 6, 5), (1, 3, 1), (0, 1, 2), (2, 1, 2), (1, 0, 2), 0),     # destination re-ordered (transposing scatter)
    ((8, 16, 6), (2, 1, 1), (1, 0, 2), (1, 2, 2), (2, 0, 1), 1),     # source re-ordered: contiguous along dim 1
    ((5, 7, 64), (1, 1, 1), (0, 1, 2), (1, 7, 1), (0, 1, 2), 2),     # single source rank, seven destinations
    ((16, 4, 4), (1, 2, 2), (0, 1, 2), (1, 2, 2), (2, 1, 0), 0),     # same extents, new order: a local permutation through the map
    ((20, 21, 22), (1, 1, 2), (0, 1, 2), (1, 2, 1), (0, 1, 2), 1),   # generic kernel along the middle axis, uneven halves
    ((10, 9, 22), (1, 2, 1), (0, 1, 2), (1, 1, 2), (0, 1, 2), 2),    # generic kernel along the slow axis
    ((7, 6, 5), (1, 1, 1), (0, 1, 2), (1, 1, 1), (0, 1, 2), 1),      # one rank, one cell
    # whole tiles per box row and destination ranges that divide the slow line axis: the tiles visit the ranges round-robin
    ((64, 16, 8), (1, 1, 1), (0, 1, 2), (2, 1, 2), (0, 1, 2), 1),    # strided kernel, 2 (fp64) / 1 (fp32) tiles per row, nb = 2
    ((16, 64, 6), (1, 1, 1), (0, 1, 2), (1, 2, 3), (0, 1, 2), 0),    # contig kernel, nb = 3
    ((64, 4, 16), (1, 2, 1), (0, 1, 2), (2, 4, 1), (0, 1, 2), 2),    # strided kernel along the slow axis, nb = 2 per source rank
    # lengths with factors 3 and 5 (mixed-radix fast kernels), fused reshape
    ((192, 6, 4), (1, 2, 2), (0, 1, 2), (3, 1, 2), (0, 1, 2), 0),    # contig kernel, n = 192 = 8 * 8 * 3
    ((5, 96, 3), (1, 1, 3), (0, 1, 2), (1, 2, 1), (0, 1, 2), 1),     # strided kernel, n = 96 = 12 * 8
    ((4, 3, 200), (2, 1, 1), (0, 1, 2), (1, 1, 4), (0, 1, 2), 2),    # strided kernel along the slow axis, n = 200 = 10 * 10 * 2
    ((80, 8, 2), (1, 1, 2), (0, 1, 2), (5, 2, 1), (0, 1, 2), 0),     # contig kernel, n = 80 = 5 * 4 * 4
]


@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_fft_with_fused_reshape(emul, prec, case):  # noqa: F811
    from heffte_b200 import _lib
    n, gsrc, osrc, gdst, odst, dim = case
    world = O.world_box(n)
    src_boxes, dst_boxes = bricks(world, gsrc, osrc), bricks(world, gdst, odst)
    rng = np.random.default_rng(42)
    x = (rng.random(world.count()) + 1j * rng.random(world.count()))
    cdt = np.complex64 if prec == 0 else np.complex128
    for direction in (0, 1):
        expect_world = O.exec1d_c2c(x, world, dim, backward=bool(direction)) * 0.25
        outs = [np.zeros(max(b.count(), 1), dtype=cdt) for b in dst_boxes]
        for me, box in enumerate(src_boxes):
            if box.count() == 0:
                continue
            local = np.ascontiguousarray(O.get_subbox(world, box, x).astype(cdt))
            g, ca, cb = unlumped_geometry(box, dim)
            d = _lib.b200_fft1d_desc(prec, 0, box.size[dim], ca, cb, _lib.b200_line_geom(*g), _lib.b200_line_geom(*g))
            mine = _nine([box])
            rc = emul.emul_fft1d_reshape(ctypes.byref(d), direction, local.ctypes.data, mine.ctypes.data, dim, len(dst_boxes),
                                         _nine(dst_boxes).ctypes.data, _bases(outs), local.itemsize, ctypes.c_double(0.25))
            assert rc == 0
        for b, got in zip(dst_boxes, outs):
            if b.count():
                assert O.rel_l2(got[:b.count()], O.get_subbox(world, b, expect_world)) < (3e-6 if prec == 0 else 1e-13)


@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("n,dim,src_grid,dst_grid", [
    ((12, 6, 5), 0, (1, 2, 2), (3, 1, 2)),      # generic kernel
    ((64, 6, 5), 0, (1, 2, 2), (3, 1, 2)),      # fft_contig_real_kernel, scatter variant
    ((6, 64, 4), 1, (2, 1, 2), (1, 3, 2)),      # fft_strided_real_kernel (middle axis), scatter variant
    ((5, 3, 32), 2, (2, 2, 1), (1, 1, 3)),      # fft_strided_real_kernel (slow axis), scatter variant
    ((40, 128, 1), 1, (3, 1, 1), (2, 2, 1)),    # more lines than one tile row, ragged last tile
    ((32, 64, 4), 0, (1, 1, 1), (1, 2, 2)),     # contiguous real kernel, whole tiles per row: round-robin visit of nb = 2 ranges
    ((64, 32, 4), 1, (1, 1, 1), (2, 1, 2)),     # strided real kernel, whole tiles per row: round-robin visit of nb = 2 ranges
    ((160, 4, 3), 0, (1, 2, 1), (2, 1, 3)),     # contiguous real kernel on the mixed-radix engine (m = 80 = 4 * 4 * 5)
    ((6, 192, 2), 1, (2, 1, 1), (1, 3, 2)),     # strided real kernel on the mixed-radix engine (m = 96 = 12 * 8)
])
@pytest.mark.parametrize("kind", ["r2c", "c2r", "cos", "sin", "cos_b", "sin_b"])
def test_real_transforms_with_fused_reshape(emul, kind, n, dim, src_grid, dst_grid, prec):  # noqa: F811
    from heffte_b200 import _lib
    world = O.world_box(n)
    cworld = world.r2c(dim)
    rng = np.random.default_rng(1)
    rdt, cdt = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    if kind == "r2c":
        x = rng.random(world.count())
        src_boxes, dst_world = bricks(world, src_grid), cworld
        expect_world = O.exec1d_r2c(x, world, dim)
        in_world, in_dtype, out_dtype, kcode, direction = world, rdt, cdt, 1, 0
    elif kind == "c2r":
        x = O.exec1d_r2c(rng.random(world.count()), world, dim)
        src_boxes, dst_world = bricks(cworld, src_grid), world
        expect_world = O.exec1d_c2r(x, world, dim)
        in_world, in_dtype, out_dtype, kcode, direction = cworld, cdt, rdt, 1, 1
    else:
        name, backward = kind[:3], kind.endswith("_b")
        x = rng.random(world.count())
        src_boxes, dst_world = bricks(world, src_grid), world
        expect_world = O.r2r_backward(x, world, dim, name) if backward else O.r2r_forward(x, world, dim, name)
        in_world, in_dtype, out_dtype, kcode, direction = world, rdt, rdt, {"cos": 2, "sin": 3}[name], int(backward)
    dst_boxes = bricks(dst_world, dst_grid)
    outs = [np.zeros(max(b.count(), 1), dtype=out_dtype) for b in dst_boxes]
    for me, box in enumerate(src_boxes):
        local = np.ascontiguousarray(O.get_subbox(in_world, box, x).astype(in_dtype))
        # the descriptor always speaks about the REAL box; the scatter map about the box the kernel writes
        high = list(box.high)
        high[dim] = world.high[dim]
        rbox = O.Box(box.low, tuple(high), box.order) if kind == "c2r" else box
        cbox = rbox.r2c(dim)
        gi, ca, cb = unlumped_geometry(rbox, dim)
        go = unlumped_geometry(cbox, dim)[0] if kcode == 1 else gi
        d = _lib.b200_fft1d_desc(prec, kcode, rbox.size[dim], ca, cb, _lib.b200_line_geom(*gi), _lib.b200_line_geom(*go))
        written = cbox if kind == "r2c" else rbox
        rc = emul.emul_fft1d_reshape(ctypes.byref(d), direction, local.ctypes.data, _nine([written]).ctypes.data, dim, len(dst_boxes),
                                     _nine(dst_boxes).ctypes.data, _bases(outs), np.dtype(out_dtype).itemsize, ctypes.c_double(1.0))
        assert rc == 0
    for b, got in zip(dst_boxes, outs):
        assert O.rel_l2(got[:b.count()], O.get_subbox(dst_world, b, expect_world)) < (1e-5 if prec == 0 else 1e-12)


@pytest.mark.parametrize("elem", [np.float32, np.float64, np.complex128])
def test_scatter_copy(emul, elem):  # noqa: F811
    world = O.world_box((9, 10, 11))
    rng = np.random.default_rng(8)
    x = rng.random(world.count()).astype(elem)
    for gsrc, osrc, gdst, odst in [((1, 2, 3), (0, 1, 2), (6, 1, 1), (0, 1, 2)), ((3, 2, 1), (0, 1, 2), (1, 1, 6), (1, 2, 0)),
                                   ((2, 3, 1), (2, 0, 1), (1, 6, 1), (0, 1, 2)), ((1, 1, 1), (0, 1, 2), (2, 2, 2), (0, 1, 2))]:
        src_boxes, dst_boxes = bricks(world, gsrc, osrc), bricks(world, gdst, odst)
        outs = [np.zeros(b.count(), dtype=elem) for b in dst_boxes]
        for box in src_boxes:
            local = np.ascontiguousarray(O.get_subbox(world, box, x))
            rc = emul.emul_scatter_copy(local.itemsize, local.ctypes.data, _nine([box]).ctypes.data, len(dst_boxes), _nine(dst_boxes).ctypes.data, _bases(outs))
            assert rc == 0
        for b, got in zip(dst_boxes, outs):
            assert np.array_equal(got, O.get_subbox(world, b, x))
