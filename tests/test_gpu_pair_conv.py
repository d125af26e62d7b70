"""
GPU parity of the paired kernel (two local transforms of a box in one persistent launch, b200_fft1d_execute_pair) and of the
fused spectral-operator kernel (forward, product, backward in one pass, b200_fft1d_execute_convolve) through the C ABI, against
numpy in double precision; then the plan-level forms: batched transforms in one launch per stage and plan.convolve() against the
oracle's forward -> product -> backward (the semantics of the reference's benchmarks/convolution.cpp:86-97).
"""
import ctypes

import numpy as np
import pytest

from oracle import heffte_oracle as O
from tests.helpers import TOL, to_h

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
vp = ctypes.c_void_p


def _plan(lib, prec, n, ca, cb, geom):
    from heffte_b200._lib import b200_fft1d_desc, b200_line_geom
    d = b200_fft1d_desc(prec, 0, n, ca, cb, b200_line_geom(*geom), b200_line_geom(*geom))
    p = vp()
    assert lib.b200_fft1d_create(ctypes.byref(d), ctypes.byref(p)) == 0
    return p


@pytest.mark.parametrize("prec,n,planes", [(1, 128, 5), (0, 128, 7), (1, 256, 9), (0, 256, 33), (1, 512, 6), (0, 512, 5), (1, 1024, 2), (0, 1024, 3)])
def test_paired_kernel(lib, prec, n, planes):
    cdt = np.complex64 if prec == 0 else np.complex128
    tol = TOL[prec]
    pc = _plan(lib, prec, n, n, planes, (1, n, n * n))
    ps = _plan(lib, prec, n, n, planes, (n, 1, n * n))
    assert lib.b200_fft1d_pairable(pc, ps) == 1
    batch = 2
    count = n * n * planes
    rng = np.random.default_rng(n + planes)
    x = (rng.random(batch * count) + 1j * rng.random(batch * count)).astype(cdt)
    counters = torch.zeros(batch * planes + 8, dtype=torch.int32, device="cuda")
    step = count * x.itemsize
    for direction in (0, 1):
        ref = x.reshape(batch, planes, n, n).astype(np.complex128)
        ref = np.fft.fft2(ref, axes=(2, 3)) if direction == 0 else np.fft.ifft2(ref, axes=(2, 3)) * (n * n)
        ref = ref.reshape(-1) * 0.25
        for first, second in ((pc, ps), (ps, pc)):
            for lag in (1, 4):
                src = torch.from_numpy(x).cuda()
                mid = torch.zeros_like(src)
                rc = lib.b200_fft1d_execute_pair(first, second, direction, vp(src.data_ptr()), vp(mid.data_ptr()), None, ctypes.c_double(0.25),
                                                 vp(counters.data_ptr()), lag, None, batch, step, step, 0, 0, 0)
                assert rc == 0
                torch.cuda.synchronize()
                assert O.rel_l2(mid.cpu().numpy(), ref) <= tol
                # in place
                rc = lib.b200_fft1d_execute_pair(first, second, direction, vp(src.data_ptr()), vp(src.data_ptr()), None, ctypes.c_double(0.25),
                                                 vp(counters.data_ptr()), lag, None, batch, step, step, 0, 0, 0)
                assert rc == 0
                torch.cuda.synchronize()
                assert O.rel_l2(src.cpu().numpy(), ref) <= tol
    lib.b200_fft1d_destroy(pc)
    lib.b200_fft1d_destroy(ps)


@pytest.mark.parametrize("prec,n,lines", [(1, 16, 1000), (0, 32, 513), (1, 64, 777), (0, 128, 300), (1, 256, 129), (0, 512, 1000), (1, 512, 4096),
                                          (0, 1024, 64), (1, 2048, 40), (0, 4096, 33), (1, 4096, 8)])
def test_spectral_operator_kernel(lib, prec, n, lines):
    cdt = np.complex64 if prec == 0 else np.complex128
    tol = 4 * TOL[prec]
    p = _plan(lib, prec, n, lines, 1, (lines, 1, 0))
    assert lib.b200_fft1d_convolvable(p) == 1
    batch = 2
    count = n * lines
    rng = np.random.default_rng(n)
    x = (rng.random(batch * count) + 1j * rng.random(batch * count)).astype(cdt)
    m = (rng.random(count) + 1j * rng.random(count)).astype(cdt)
    scale = 1.0 / n
    spec = np.fft.fft(x.reshape(batch, n, lines).astype(np.complex128), axis=1) * scale
    step = count * x.itemsize
    dm = torch.from_numpy(m).cuda()
    for mult in (None, dm):
        prod = spec * (spec if mult is None else m.reshape(1, n, lines).astype(np.complex128))
        ref = (np.fft.ifft(prod, axis=1) * n).reshape(-1)
        src = torch.from_numpy(x).cuda()
        out = torch.zeros_like(src)
        rc = lib.b200_fft1d_execute_convolve(p, vp(src.data_ptr()), vp(out.data_ptr()), None, None if mult is None else vp(mult.data_ptr()),
                                             ctypes.c_double(scale), None, batch, step, step, 0, 0, 0)
        assert rc == 0
        torch.cuda.synchronize()
        assert O.rel_l2(out.cpu().numpy(), ref) <= tol
        rc = lib.b200_fft1d_execute_convolve(p, vp(src.data_ptr()), vp(src.data_ptr()), None, None if mult is None else vp(mult.data_ptr()),
                                             ctypes.c_double(scale), None, batch, step, step, 0, 0, 0)
        assert rc == 0
        torch.cuda.synchronize()
        assert O.rel_l2(src.cpu().numpy(), ref) <= tol
    lib.b200_fft1d_destroy(p)


@pytest.mark.parametrize("kind,n,prec,batch", [("c2c", (64, 32, 48), 1, 5), ("c2c", (128, 128, 16), 0, 3), ("c2c", (256, 256, 8), 0, 2),
                                               ("r2c", (64, 40, 24), 1, 4), ("cos", (32, 64, 16), 1, 3), ("c2c", (12, 10, 9), 0, 5)])
def test_batched_plan_single_rank(lib, kind, n, prec, batch):
    """forward(batch, ...) / backward(batch, ...): one launch per stage for all entries (reference include/heffte_fft3d.h:391-414)"""
    import heffte_b200 as hf
    world = O.world_box(n)
    rdt, cdt = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    rng = np.random.default_rng(3)
    count = world.count()
    if kind == "r2c":
        fft = hf.fft3d_r2c(hf.backend.b200, to_h(world), to_h(world.r2c(0)), 0, hf.comm_self())
    else:
        fft = hf.fft3d({"c2c": hf.backend.b200, "cos": hf.backend.b200_cos}[kind], to_h(world), to_h(world), hf.comm_self())
    xs = [rng.random(count) + (1j * rng.random(count) if kind == "c2c" else 0) for _ in range(batch)]
    x = np.concatenate(xs).astype(cdt if kind == "c2c" else rdt)
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty(batch * fft.size_outbox(), dtype=torch.from_numpy(np.zeros(1, dtype=rdt if kind == "cos" else cdt)).dtype, device="cuda")
    fft.forward(dx, dy, hf.scale.full, batch=batch)
    got = dy.cpu().numpy()
    tol = TOL[prec] * (4 if kind == "cos" else 1)
    for b in range(batch):
        ref = O.fft3d_forward(xs[b], n, kind, scaling="full")
        assert O.rel_l2(got[b * fft.size_outbox():(b + 1) * fft.size_outbox()], ref) <= tol
    dz = torch.empty_like(dx)
    fft.backward(dy, dz, hf.scale.none, batch=batch)
    assert O.rel_l2(dz.cpu().numpy(), x) <= 2 * tol


@pytest.mark.parametrize("n,prec", [((64, 64, 64), 1), ((32, 48, 128), 0), ((20, 12, 18), 1), ((128, 128, 128), 0)])
def test_convolve_single_rank(lib, n, prec):
    import heffte_b200 as hf
    world = O.world_box(n)
    cdt = np.complex64 if prec == 0 else np.complex128
    rng = np.random.default_rng(9)
    count = world.count()
    x = (rng.random(count) + 1j * rng.random(count)).astype(cdt)
    m = (rng.random(count) + 1j * rng.random(count)).astype(cdt)
    fft = hf.fft3d(hf.backend.b200, to_h(world), to_h(world), hf.comm_self())
    spectrum = O.fft3d_forward(x, n, "c2c", scaling="full")
    lo, hi, order = fft.convolve_box()
    assert lo == [0, 0, 0] and hi == [n[0] - 1, n[1] - 1, n[2] - 1]
    for mult in (None, m):
        ref = O.fft3d_backward(spectrum * (spectrum if mult is None else mult), n, "c2c", scaling="none")
        dx = torch.from_numpy(x).cuda()
        dout = torch.empty_like(dx)
        fft.convolve(dx, dout, None if mult is None else torch.from_numpy(mult).cuda(), hf.scale.full)
        assert O.rel_l2(dout.cpu().numpy(), ref) <= 4 * TOL[prec]
        fft.convolve(dx, dx, None if mult is None else torch.from_numpy(mult).cuda(), hf.scale.full)     # in place
        assert O.rel_l2(dx.cpu().numpy(), ref) <= 4 * TOL[prec]
