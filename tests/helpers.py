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
