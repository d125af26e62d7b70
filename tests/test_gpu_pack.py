"""
GPU parity of the packers / scaling / conversion kernels against the oracle restatement of include/heffte_pack3d.h,
bit-exact (pure data movement).  Mirrors test/test_units_nompi.cpp:577-637 (local transposes) and :71-85 (scaling).
"""
import ctypes
import itertools

import numpy as np
import pytest

from oracle import heffte_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
def test_direct_pack_unpack(lib, dtype):
    rng = np.random.default_rng(0)
    box = (37, 21, 13)
    data = rng.random(box[0] * box[1] * box[2]).astype(dtype)
    plan = dict(size=(17, 9, 5), line_stride=box[0], plane_stride=box[0] * box[1])
    offset = 3 * box[0] * box[1] + 4 * box[0] + 6
    packed_ref = O.direct_pack(plan, data, offset)
    d = _dev(data)
    buf = torch.zeros(packed_ref.size, dtype=d.dtype, device="cuda")
    rc = lib.b200_direct_pack(data.itemsize, 17, 9, 5, box[0], box[0] * box[1], ctypes.c_void_p(d.data_ptr() + offset * data.itemsize),
                              ctypes.c_void_p(buf.data_ptr()), None)
    assert rc == 0
    assert np.array_equal(buf.cpu().numpy(), packed_ref)
    target_ref = np.zeros_like(data)
    O.direct_unpack(plan, packed_ref, target_ref, offset)
    target = torch.zeros_like(d)
    rc = lib.b200_direct_unpack(data.itemsize, 17, 9, 5, box[0], box[0] * box[1], ctypes.c_void_p(buf.data_ptr()),
                                ctypes.c_void_p(target.data_ptr() + offset * data.itemsize), None)
    assert rc == 0
    assert np.array_equal(target.cpu().numpy(), target_ref)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
@pytest.mark.parametrize("src_order,dst_order", [p for p in itertools.product(itertools.permutations(range(3)), repeat=2) if p[0] != p[1]][::3])
def test_transpose_unpack(lib, dtype, src_order, dst_order):
    rng = np.random.default_rng(1)
    src = O.Box((0, 0, 0), (40, 34, 9), src_order)
    dst = O.Box((0, 0, 0), (40, 34, 9), dst_order)
    data = rng.random(src.count()).astype(dtype)
    entry = O.overlap_map(0, 1, dst, [src], receive=True)[0]
    p = entry["plan"]
    out_ref = np.zeros(dst.count(), dtype=dtype)
    O.transpose_unpack(p, data, out_ref, entry["offset"])
    d = _dev(data)
    out = torch.zeros(dst.count(), dtype=d.dtype, device="cuda")
    rc = lib.b200_transpose_unpack(data.itemsize, *p["size"], p["line_stride"], p["plane_stride"], p["buff_line_stride"], p["buff_plane_stride"],
                                   *p["map"], ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(out.data_ptr()), None)
    assert rc == 0, lib.b200_last_error()
    assert np.array_equal(out.cpu().numpy(), out_ref)


def test_reference_transpose_golden(lib):
    # test/test_units_nompi.cpp:595-636: 2x3x4 box with entries 1..24, orders (1,2,0), (2,1,0), (0,2,1)
    data = np.arange(1.0, 25.0)
    src = O.Box((0, 0, 0), (1, 2, 3))
    golden = {
        (1, 2, 0): [1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24],
        (2, 1, 0): [1, 7, 13, 19, 3, 9, 15, 21, 5, 11, 17, 23, 2, 8, 14, 20, 4, 10, 16, 22, 6, 12, 18, 24],
        (0, 2, 1): [1, 2, 7, 8, 13, 14, 19, 20, 3, 4, 9, 10, 15, 16, 21, 22, 5, 6, 11, 12, 17, 18, 23, 24],
    }
    d = _dev(data)
    for order, expected in golden.items():
        dst = O.Box((0, 0, 0), (1, 2, 3), order)
        p = O.overlap_map(0, 1, dst, [src], receive=True)[0]["plan"]
        out = torch.zeros(24, dtype=torch.float64, device="cuda")
        rc = lib.b200_transpose_unpack(8, *p["size"], p["line_stride"], p["plane_stride"], p["buff_line_stride"], p["buff_plane_stride"],
                                       *p["map"], ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(out.data_ptr()), None)
        assert rc == 0
        assert out.cpu().numpy().tolist() == [float(v) for v in expected]


def test_scale_and_convert(lib):
    rng = np.random.default_rng(2)
    for prec, rt, ct in [(0, np.float32, np.complex64), (1, np.float64, np.complex128)]:
        x = rng.random(100003).astype(rt)
        d = _dev(x)
        assert lib.b200_scale(prec, x.size, ctypes.c_void_p(d.data_ptr()), ctypes.c_double(0.37), None) == 0
        assert np.array_equal(d.cpu().numpy(), x * rt(0.37))
        d = _dev(x)
        c = torch.zeros(x.size, dtype=torch.complex64 if prec == 0 else torch.complex128, device="cuda")
        assert lib.b200_convert_r2c(prec, x.size, ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(c.data_ptr()), None) == 0
        assert np.array_equal(c.cpu().numpy(), x.astype(ct))
        z = (rng.random(5001) + 1j * rng.random(5001)).astype(ct)
        dz = _dev(z)
        r = torch.zeros(z.size, dtype=torch.float32 if prec == 0 else torch.float64, device="cuda")
        assert lib.b200_convert_c2r(prec, z.size, ctypes.c_void_p(dz.data_ptr()), ctypes.c_void_p(r.data_ptr()), None) == 0
        assert np.array_equal(r.cpu().numpy(), z.real)
