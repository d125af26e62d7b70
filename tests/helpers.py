"""Shared helpers of the parity tests."""
import numpy as np

from oracle import heffte_oracle as O

# tolerances of the task statement: relative L2 vs the reference result
TOL = {0: 1e-5, 1: 1e-12}


def seeded(count, seed, complex_values):
    rng = np.random.default_rng(seed)
    x = rng.random(count)
    if complex_values:
        x = x + 1j * rng.random(count)
    return x


def to_h(box):
    from heffte_b200 import heffte as H
    return H.box3d(box.low, box.high, box.order)


def bricks(world, grid, order=(0, 1, 2)):
    """near-equal split of the world (same rule as heffte::split_world), pure python so it works without any library"""
    n = world.size
    cut = lambda d, i: world.low[d] + i * (n[d] // grid[d]) + min(i, n[d] % grid[d])
    out = []
    for k in range(grid[2]):
        for j in range(grid[1]):
            for i in range(grid[0]):
                out.append(O.Box((cut(0, i), cut(1, j), cut(2, k)), (cut(0, i + 1) - 1, cut(1, j + 1) - 1, cut(2, k + 1) - 1), order))
    return out


def line_geometry(box, dim):
    """(stride, stride_a, stride_b), count_a, count_b of the lines of `box` along `dim`"""
    o = box.order
    if dim == o[0]:
        return (1, box.osize(0), 0), box.osize(1) * box.osize(2), 1
    if dim == o[1]:
        return (box.osize(0), 1, box.osize(0) * box.osize(1)), box.osize(0), box.osize(2)
    return (box.osize(0) * box.osize(1), 1, 0), box.osize(0) * box.osize(1), 1


def host_ranks(limit=16):
    """power-of-two number of thread-ranks for the compiled reference: as many as the host has cores, at most `limit`"""
    import os
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    r = 1
    while r * 2 <= min(cores, limit):
        r *= 2
    return r


def reference_world_forward(ref, kind, prec, n, x, scaling="none", r2c_dir=0, ranks=None):
    """
    The WHOLE transform of the world array x (flat, order (0,1,2)) by the compiled reference (oracle/_ref, stock backend)
    spread over thread-ranks as slabs of the slowest axis -- the test-suite pattern of the reference itself
    (test/test_fft3d.h:124-214: world array -> sub-boxes -> fft3d -> compare).  Returns the flat world output.
    """
    ranks = ranks or host_ranks()
    ranks = max(1, min(ranks, n[2]))
    world = O.world_box(n)
    out_world = world.r2c(r2c_dir) if kind == "r2c" else world
    inboxes, outboxes = bricks(world, (1, 1, ranks)), bricks(out_world, (1, 1, ranks))
    plane_in, plane_out = world.size[0] * world.size[1], out_world.size[0] * out_world.size[1]
    inputs = [x[b.low[2] * plane_in:(b.high[2] + 1) * plane_in] for b in inboxes]
    outs, _ = ref.fft3d(kind, prec, inboxes, outboxes, inputs, scaling=scaling, r2c_dir=r2c_dir)
    assert all(o.size == b.count() for o, b in zip(outs, outboxes)) and plane_out > 0
    return np.concatenate(outs)
