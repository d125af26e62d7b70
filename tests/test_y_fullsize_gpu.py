"""
The BASELINE problem sizes on one GPU.  First against the reference itself: oracle/_ref (the unmodified reference, stock
backend, thread-ranks) transforms the same seeded world array in a few seconds and the outputs are compared in relative L2
(<= 1e-12 fp64, <= 1e-5 fp32) -- the pattern of the reference's own test/test_fft3d.h:124-214.  Then through
size-independent properties: Parseval, linearity, the closed-form spectrum of a shifted delta, the round trip with
scale::full, and the agreement between the r2c plan and the complex plan on the half spectrum.  Sizes: 512^3 fp64 (bench line, r2c, DCT) and
256^3 fp32 (speed3d_c2c single 256^3 of BASELINE.json).  The file sorts late on purpose: it is the heaviest of the GPU suite.
"""
import math

import numpy as np
import pytest

from oracle import heffte_oracle as O
from tests.helpers import TOL, reference_world_forward, to_h

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

FULL = (512, 512, 512)      # speed3d_c2c / r2c / r2r double 512^3 of BASELINE.json


def _rel(a, b):
    return float((torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item())


def _plan(kind, n, hf):
    world = O.world_box(n)
    if kind == "r2c":
        return hf.fft3d_r2c(hf.backend.b200, to_h(world), to_h(world.r2c(0)), 0, hf.comm_self())
    tag = {"c2c": hf.backend.b200, "cos": hf.backend.b200_cos, "sin": hf.backend.b200_sin}[kind]
    return hf.fft3d(tag, to_h(world), to_h(world), hf.comm_self())


@pytest.mark.parametrize("n,prec", [((512, 512, 512), 1), ((256, 256, 256), 0)])
def test_c2c_properties_at_full_size(lib, n, prec):
    import heffte_b200 as hf
    tol = TOL[prec]
    rt, ct = (torch.float32, torch.complex64) if prec == 0 else (torch.float64, torch.complex128)
    count = n[0] * n[1] * n[2]
    fft = _plan("c2c", n, hf)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4242)
    x1 = torch.complex(torch.rand(count, dtype=rt, device="cuda", generator=gen), torch.rand(count, dtype=rt, device="cuda", generator=gen))
    x2 = torch.complex(torch.rand(count, dtype=rt, device="cuda", generator=gen), torch.rand(count, dtype=rt, device="cuda", generator=gen))
    y1, y2 = torch.empty_like(x1), torch.empty_like(x2)
    fft.forward(x1, y1)
    fft.forward(x2, y2)
    # Parseval: ||F x||^2 = N ||x||^2
    energy_in = float(torch.linalg.vector_norm(x1).item()) ** 2 * count
    energy_out = float(torch.linalg.vector_norm(y1).item()) ** 2
    assert abs(energy_out - energy_in) <= 10 * tol * energy_in
    # linearity: F(a x1 + x2) = a F(x1) + F(x2)
    a = 0.375 - 1.25j
    combo = torch.empty_like(x1)
    fft.forward(a * x1 + x2, combo)
    assert _rel(combo, a * y1 + y2) <= tol
    # round trip with scale::full, and in place with a caller workspace like speed3d (benchmarks/speed3d.h:163-177)
    back = torch.empty_like(x1)
    fft.backward(y1, back, hf.scale.full)
    assert _rel(back, x1) <= tol
    work = torch.empty(fft.size_workspace(), dtype=ct, device="cuda")
    data = x2.clone()
    fft.forward_buffered(data, data, work, hf.scale.full)
    fft.backward_buffered(data, data, work, hf.scale.none)
    assert _rel(data, x2) <= tol
    # closed form: a delta at (1, 2, 3) transforms into the plane wave exp(-2 pi i (k0 + 2 k1 + 3 k2) / n)
    del x2, y2, combo, back, data
    delta = torch.zeros(count, dtype=ct, device="cuda")
    delta[1 + n[0] * (2 + n[1] * 3)] = 1.0
    fft.forward(delta, y1)
    rng = np.random.default_rng(7)
    k = rng.integers(0, [n[0], n[1], n[2]], size=(4096, 3))
    index = torch.from_numpy(k[:, 0] + n[0] * (k[:, 1] + n[1] * k[:, 2])).cuda()
    phase = -2.0 * math.pi * (k[:, 0] * 1.0 / n[0] + k[:, 1] * 2.0 / n[1] + k[:, 2] * 3.0 / n[2])
    expect = torch.from_numpy(np.exp(1j * phase)).to(ct).cuda()
    assert _rel(y1[index], expect) <= tol
    assert abs(float(torch.linalg.vector_norm(y1).item()) ** 2 - count) <= 10 * tol * count


def test_r2c_matches_the_complex_plan_at_512(lib):
    """speed3d_r2c double 512^3: the half spectrum equals the first 257 entries (dimension 0) of the complex transform"""
    import heffte_b200 as hf
    n = FULL
    count, half = n[0] * n[1] * n[2], (n[0] // 2 + 1) * n[1] * n[2]
    gen = torch.Generator(device="cuda")
    gen.manual_seed(11)
    x = torch.rand(count, dtype=torch.float64, device="cuda", generator=gen)
    rfft, cfft = _plan("r2c", n, hf), _plan("c2c", n, hf)
    assert rfft.size_inbox() == count and rfft.size_outbox() == half
    yr = torch.empty(half, dtype=torch.complex128, device="cuda")
    rfft.forward(x, yr)
    yc = torch.empty(count, dtype=torch.complex128, device="cuda")
    cfft.forward(torch.complex(x, torch.zeros_like(x)), yc)
    expect = yc.reshape(n[2], n[1], n[0])[:, :, : n[0] // 2 + 1].reshape(-1)
    assert _rel(yr, expect) <= TOL[1]
    del yc, expect
    back = torch.empty_like(x)
    rfft.backward(yr, back, hf.scale.full)
    assert _rel(back, x) <= TOL[1]


def test_dct_properties_at_512(lib):
    """speed3d_r2r double 512^3 (DCT-II forward, DCT-III backward): round trip with scale::full, linearity, and the
    zero-frequency entry 8 * sum(x) of the unnormalised REDFT10 (include/heffte_fft3d.h:737-746)"""
    import heffte_b200 as hf
    n = FULL
    count = n[0] * n[1] * n[2]
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    x1 = torch.rand(count, dtype=torch.float64, device="cuda", generator=gen)
    x2 = torch.rand(count, dtype=torch.float64, device="cuda", generator=gen)
    fft = _plan("cos", n, hf)
    y1, y2, combo = torch.empty_like(x1), torch.empty_like(x1), torch.empty_like(x1)
    fft.forward(x1, y1)
    fft.forward(x2, y2)
    fft.forward(2.5 * x1 - x2, combo)
    assert _rel(combo, 2.5 * y1 - y2) <= 4 * TOL[1]
    assert abs(float(y1[0].item()) - 8.0 * float(x1.sum().item())) <= 1e-10 * 8.0 * float(x1.sum().item())
    back = torch.empty_like(x1)
    fft.forward(x1, y1, hf.scale.full)
    fft.backward(y1, back)
    assert _rel(back, x1) <= 4 * TOL[1]


# ---- the BASELINE sizes against the compiled reference (oracle/_ref) ------------------------------------------------------
@pytest.mark.parametrize("kind,n,prec", [("c2c", (512, 512, 512), 1), ("r2c", (512, 512, 512), 1), ("cos", (512, 512, 512), 1),
                                         ("c2c", (256, 256, 256), 0), ("sin", (256, 256, 256), 1), ("r2c", (256, 256, 256), 0)])
def test_full_size_against_the_compiled_reference(lib, reference, kind, n, prec):
    """forward(scale::full) of the seeded world array: b200 on the GPU vs heffte::fft3d<stock> (fp64) on the host cores"""
    import heffte_b200 as hf
    count = n[0] * n[1] * n[2]
    rng = np.random.default_rng(4242)
    rt, ct = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    real_in = kind != "c2c"
    x = rng.random(count).astype(rt)
    if not real_in:
        x = (x + 1j * rng.random(count).astype(rt)).astype(ct)
    # the fp32 truth is the fp64 reference result of the same (fp32-representable) input, SURVEY 8(c)
    expect = reference_world_forward(reference, kind, 1, n, x.astype(np.float64 if real_in else np.complex128), scaling="full")
    fft = _plan(kind, n, hf)
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty(fft.size_outbox(), dtype=torch.from_numpy(np.zeros(1, dtype=rt if kind in ("cos", "sin") else ct)).dtype, device="cuda")
    fft.forward(dx, dy, hf.scale.full)
    got = dy.cpu().numpy()
    del dy, dx
    assert got.shape == expect.shape
    err = float(np.linalg.norm(got.astype(expect.dtype) - expect) / np.linalg.norm(expect))
    tol = TOL[prec] * (4 if kind in ("cos", "sin") else 1)
    assert err <= tol, (kind, n, prec, err)
    if prec == 0 and kind == "c2c":
        # for the record: the distance of the reference's own fp32 (stock) result from the same truth
        stock32 = reference_world_forward(reference, kind, 0, n, x, scaling="full")
        print("fp32 %s %s: b200 %.3e, stock fp32 %.3e (both vs the fp64 reference)" % (
            kind, n, err, float(np.linalg.norm(stock32.astype(expect.dtype) - expect) / np.linalg.norm(expect))))
