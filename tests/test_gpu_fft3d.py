"""
GPU parity of the whole transform on one rank through the reference-facing plan API (heffte_plan_create /
heffte_forward_* in include/heffte_b200.h, via the Python binding that mirrors python/heffte.py of the reference).
Mirrors test/test_fft3d_np1.cpp / test/test_fft3d.h:160-250 (all option combinations x 3 scalings, c2c and real input,
in-place) and test/test_fft3d_r2c.cpp (three r2c directions) and test/test_cos.cpp (r2r), with the oracle standing
where the reference uses its own three 1-D executor sweeps (test/test_fft3d.h:124-155).
"""
import numpy as np
import pytest

from oracle import heffte_oracle as O
from tests.helpers import TOL, to_h

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

SCALINGS = ["none", "full", "symmetric"]
SIZES = [(4, 4, 4), (6, 7, 8), (17, 16, 16), (64, 64, 64), (21, 20, 19), (32, 1, 16), (128, 64, 32)]


def _data(count, prec, complex_values, ref=None):
    # the reference's make_data (minstd_rand(4242) -> U(0,1)) when oracle/_ref is present, else a seeded numpy stream
    from oracle import ref_lib
    x = ref_lib.make_data(count) if ref_lib.available() else np.random.default_rng(4242).random(count)
    if complex_values:
        return x.astype(np.complex64 if prec == 0 else np.complex128)
    return x.astype(np.float32 if prec == 0 else np.float64)


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("reorder", [False, True])
def test_c2c_single_rank(lib, prec, n, reorder):
    import heffte_b200 as hf
    world = O.world_box(n)
    comm = hf.comm_self()
    for pencils in (True, False):
        fft = hf.fft3d(hf.backend.b200, to_h(world), to_h(world), comm, hf.plan_options(hf.backend.b200, use_reorder=reorder, use_pencils=pencils))
        assert fft.size_inbox() == world.count() and fft.size_outbox() == world.count()
        x = _data(world.count(), prec, True)
        dx = torch.from_numpy(x).cuda()
        for si, scaling in enumerate(SCALINGS):
            dy = torch.empty_like(dx)
            fft.forward(dx, dy, si)
            ref = O.fft3d_forward(x, n, "c2c", scaling=scaling)
            assert O.rel_l2(dy.cpu().numpy(), ref) <= TOL[prec]
            dz = torch.empty_like(dx)
            fft.backward(dy, dz, si)
            refb = O.fft3d_backward(ref, n, "c2c", scaling=scaling)
            assert O.rel_l2(dz.cpu().numpy(), refb) <= TOL[prec]
        # in-place with user workspace (test/test_fft3d.h:398-440)
        work = torch.empty(fft.size_workspace(), dtype=dx.dtype, device="cuda")
        dy = dx.clone()
        fft.forward_buffered(dy, dy, work, hf.scale.none)
        assert O.rel_l2(dy.cpu().numpy(), O.fft3d_forward(x, n, "c2c")) <= TOL[prec]
        fft.backward_buffered(dy, dy, work, hf.scale.full)
        assert O.rel_l2(dy.cpu().numpy(), x) <= TOL[prec]


@pytest.mark.parametrize("prec", [0, 1])
def test_c2c_real_input_and_output(lib, prec):
    # reference API: forward(real in, complex out), backward(complex in, real out) on a c2c plan (heffte_fft3d.h:353-384)
    import heffte_b200 as hf
    n = (12, 9, 10)
    world = O.world_box(n)
    fft = hf.fft3d(hf.backend.b200, to_h(world), to_h(world), hf.comm_self())
    x = _data(world.count(), prec, False)
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty(world.count(), dtype=torch.complex64 if prec == 0 else torch.complex128, device="cuda")
    fft.forward(dx, dy)
    ref = O.fft3d_forward(x, n, "c2c")
    assert O.rel_l2(dy.cpu().numpy(), ref) <= TOL[prec]
    dz = torch.empty_like(dx)
    fft.backward(dy, dz, hf.scale.full)
    assert O.rel_l2(dz.cpu().numpy(), x) <= TOL[prec]


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("n", [(4, 4, 4), (6, 7, 8), (17, 16, 16), (64, 32, 48), (9, 10, 11)])
@pytest.mark.parametrize("r2c_dir", [0, 1, 2])
def test_r2c_single_rank(lib, prec, n, r2c_dir):
    import heffte_b200 as hf
    world = O.world_box(n)
    cworld = world.r2c(r2c_dir)
    for reorder in (False, True):
        fft = hf.fft3d_r2c(hf.backend.b200, to_h(world), to_h(cworld), r2c_dir, hf.comm_self(), hf.plan_options(hf.backend.b200, use_reorder=reorder))
        assert fft.size_inbox() == world.count() and fft.size_outbox() == cworld.count()
        x = _data(world.count(), prec, False)
        dx = torch.from_numpy(x).cuda()
        dy = torch.empty(cworld.count(), dtype=torch.complex64 if prec == 0 else torch.complex128, device="cuda")
        for si, scaling in enumerate(SCALINGS):
            fft.forward(dx, dy, si)
            ref = O.fft3d_forward(x, n, "r2c", r2c_dir=r2c_dir, scaling=scaling)
            assert O.rel_l2(dy.cpu().numpy(), ref) <= TOL[prec]
            dz = torch.empty_like(dx)
            fft.backward(dy, dz, si)
            refb = O.fft3d_backward(ref, n, "r2c", r2c_dir=r2c_dir, scaling=scaling)
            assert O.rel_l2(dz.cpu().numpy(), refb) <= 2 * TOL[prec]


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("kind", ["cos", "sin", "cos1"])
@pytest.mark.parametrize("n", [(2, 3, 4), (8, 8, 8), (5, 6, 7), (32, 16, 24)])
def test_r2r_single_rank(lib, prec, kind, n):
    import heffte_b200 as hf
    world = O.world_box(n)
    tag = {"cos": hf.backend.b200_cos, "sin": hf.backend.b200_sin, "cos1": hf.backend.b200_cos1}[kind]
    fft = hf.fft3d(tag, to_h(world), to_h(world), hf.comm_self())
    x = _data(world.count(), prec, False)
    dx = torch.from_numpy(x).cuda()
    for si, scaling in enumerate(SCALINGS):
        dy = torch.empty_like(dx)
        fft.forward(dx, dy, si)
        ref = O.fft3d_forward(x, n, kind, scaling=scaling)
        assert O.rel_l2(dy.cpu().numpy(), ref) <= 4 * TOL[prec]
        dz = torch.empty_like(dx)
        fft.backward(dy, dz, si)
        assert O.rel_l2(dz.cpu().numpy(), O.fft3d_backward(ref, n, kind, scaling=scaling)) <= 4 * TOL[prec]
    # forward(full) then backward == identity (the scale table of heffte_fft3d.h:635-650)
    dy = torch.empty_like(dx)
    fft.forward(dx, dy, hf.scale.full)
    dz = torch.empty_like(dx)
    fft.backward(dy, dz, hf.scale.none)
    assert O.rel_l2(dz.cpu().numpy(), x) <= 4 * TOL[prec]


def test_r2r_golden_test_cos(lib):
    # test/test_cos.cpp:36-54: 3-D DCT-II / DST-II / DCT-I of the 2x3x4 iota input (golden 24-vectors of the reference)
    import heffte_b200 as hf
    from tests.golden.known_answers import COS_2x3x4, SIN_2x3x4, COS1_2x3x4
    world = O.world_box((2, 3, 4))
    x = np.arange(1.0, 25.0)
    for tag, expect in [(hf.backend.b200_cos, COS_2x3x4), (hf.backend.b200_sin, SIN_2x3x4), (hf.backend.b200_cos1, COS1_2x3x4)]:
        fft = hf.fft3d(tag, to_h(world), to_h(world), hf.comm_self())
        dx = torch.from_numpy(x).cuda()
        dy = torch.empty_like(dx)
        fft.forward(dx, dy)
        assert np.allclose(dy.cpu().numpy(), expect, rtol=0, atol=1e-10 * max(1.0, np.abs(expect).max()))


def test_host_path_numpy(lib):
    # end-to-end entry (host buffers in, host buffers out): heffte_execute_host
    import heffte_b200 as hf
    n = (16, 12, 10)
    world = O.world_box(n)
    fft = hf.fft3d(hf.backend.b200, to_h(world), to_h(world), hf.comm_self())
    x = _data(world.count(), 1, True)
    y = np.empty_like(x)
    fft.forward(x, y, hf.scale.symmetric)
    assert O.rel_l2(y, O.fft3d_forward(x, n, "c2c", scaling="symmetric")) <= TOL[1]


def test_batch(lib):
    # test/test_fft3d.h:505-572 (batch of 5 transforms in one call)
    import heffte_b200 as hf
    n = (8, 9, 10)
    world = O.world_box(n)
    fft = hf.fft3d(hf.backend.b200, to_h(world), to_h(world), hf.comm_self())
    batch = 5
    x = _data(batch * world.count(), 1, True)
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty_like(dx)
    fft.forward(dx, dy, hf.scale.none, batch=batch)
    for b in range(batch):
        seg = slice(b * world.count(), (b + 1) * world.count())
        assert O.rel_l2(dy[seg].cpu().numpy(), O.fft3d_forward(x[seg], n, "c2c")) <= TOL[1]


def test_errors(lib):
    import heffte_b200 as hf
    world = O.world_box((4, 4, 4))
    with pytest.raises(hf.heffte_input_error):
        hf.fft3d(hf.backend.cufft, to_h(world), to_h(world), hf.comm_self())     # backend not provided by this library
    fft = hf.fft3d(hf.backend.b200, to_h(world), to_h(world), hf.comm_self())
    with pytest.raises(hf.heffte_input_error):
        fft.forward(torch.zeros(3, dtype=torch.complex128, device="cuda"), torch.zeros(64, dtype=torch.complex128, device="cuda"))
    with pytest.raises(hf.heffte_input_error):
        hf.fft3d_r2c(hf.backend.b200, to_h(world), to_h(world.r2c(0)), 5, hf.comm_self())
