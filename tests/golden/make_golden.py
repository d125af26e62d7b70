"""
TEST INFRASTRUCTURE.  Generates the committed golden fixtures (tests/golden/*.npz, *.json) from the UNMODIFIED reference
(stock backend, oracle/_ref/libheffte_ref.so, built by oracle/Makefile from /root/reference).  Run in the development
container only (the GPU box has no /root/reference):

    python -m tests.golden.make_golden

The fixtures are small on purpose; they pin the numpy oracle (tests/test_oracle.py), the planner
(tests/test_plan_logic.py) and, on the GPU, the CUDA path (tests/test_gpu_golden.py).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import heffte_oracle as O  # noqa: E402
from oracle import ref_lib  # noqa: E402
from tests.helpers import bricks  # noqa: E402

# (kind, n, in grid, out grid, options) -- distributed transforms run by the reference on thread-ranks
FFT3D_CASES = [
    dict(name="c2c_4x4x4_np2", kind="c2c", n=(4, 4, 4), gin=(1, 1, 2), gout=(1, 1, 2), reorder=True, pencils=True, alg=0),
    dict(name="c2c_12x10x8_np4", kind="c2c", n=(12, 10, 8), gin=(1, 2, 2), gout=(2, 2, 1), reorder=False, pencils=True, alg=0),
    dict(name="c2c_9x11x13_np3_slab", kind="c2c", n=(9, 11, 13), gin=(1, 3, 1), gout=(3, 1, 1), reorder=True, pencils=False, alg=2),
    dict(name="c2c_16x16x16_np8", kind="c2c", n=(16, 16, 16), gin=(2, 2, 2), gout=(2, 2, 2), reorder=False, pencils=True, alg=0),
    dict(name="r2c0_12x10x8_np4", kind="r2c", n=(12, 10, 8), gin=(1, 2, 2), gout=(2, 2, 1), reorder=True, pencils=True, alg=0, r2c_dir=0),
    dict(name="r2c1_8x9x10_np2", kind="r2c", n=(8, 9, 10), gin=(2, 1, 1), gout=(1, 1, 2), reorder=False, pencils=True, alg=3, r2c_dir=1),
    dict(name="r2c2_6x7x8_np2", kind="r2c", n=(6, 7, 8), gin=(1, 2, 1), gout=(2, 1, 1), reorder=True, pencils=True, alg=0, r2c_dir=2),
    dict(name="cos_6x5x4_np2", kind="cos", n=(6, 5, 4), gin=(1, 1, 2), gout=(2, 1, 1), reorder=True, pencils=True, alg=0),
    dict(name="sin_6x5x4_np2", kind="sin", n=(6, 5, 4), gin=(1, 1, 2), gout=(2, 1, 1), reorder=True, pencils=True, alg=0),
    dict(name="cos1_6x5x4_np2", kind="cos1", n=(6, 5, 4), gin=(1, 1, 2), gout=(2, 1, 1), reorder=True, pencils=True, alg=0),
]

# 1-D executor cases: (kind, box low, box high, order, dim)
EXEC1D_CASES = [
    ("c2c", (0, 0, 0), (7, 5, 3), (0, 1, 2), 0), ("c2c", (0, 0, 0), (7, 5, 3), (0, 1, 2), 1), ("c2c", (0, 0, 0), (7, 5, 3), (0, 1, 2), 2),
    ("c2c", (1, 2, 3), (6, 9, 7), (2, 0, 1), 0), ("c2c", (1, 2, 3), (6, 9, 7), (1, 2, 0), 2), ("c2c", (0, 0, 0), (15, 1, 2), (0, 1, 2), 0),
    ("r2c", (0, 0, 0), (7, 5, 3), (0, 1, 2), 0), ("r2c", (0, 0, 0), (6, 5, 3), (0, 1, 2), 1), ("r2c", (0, 0, 0), (4, 5, 8), (1, 0, 2), 2),
    ("cos", (0, 0, 0), (7, 2, 3), (0, 1, 2), 0), ("sin", (0, 0, 0), (6, 2, 3), (0, 1, 2), 0), ("cos1", (0, 0, 0), (8, 2, 3), (0, 1, 2), 0),
]

# plans of the BASELINE.json configurations and of the reference's own test geometries: (n, nranks, options ...)
PLAN_CASES = [
    dict(name="target_512_np1", n=(512, 512, 512), np=1), dict(name="target_512_np2", n=(512, 512, 512), np=2),
    dict(name="target_512_np4", n=(512, 512, 512), np=4), dict(name="target_512_np8", n=(512, 512, 512), np=8),
    dict(name="cfg2_256_np1_reorder", n=(256, 256, 256), np=1, reorder=True),
    dict(name="cfg3_r2c_512_np2", n=(512, 512, 512), np=2, r2c_dir=0), dict(name="cfg3_r2c_512_np4", n=(512, 512, 512), np=4, r2c_dir=0),
    dict(name="cfg3_r2c_512_np8", n=(512, 512, 512), np=8, r2c_dir=0),
    dict(name="cfg4_1024_np8_pencils", n=(1024, 1024, 1024), np=8, reorder=True, alg=3),
    dict(name="cfg4_1024_np8_slabs", n=(1024, 1024, 1024), np=8, reorder=True, alg=3, pencils=False),
    dict(name="cfg5_r2r_512_np8", n=(512, 512, 512), np=8, reorder=True),
    dict(name="odd_33x21x15_np6", n=(33, 21, 15), np=6, reorder=False), dict(name="odd_33x21x15_np12_slabs", n=(33, 21, 15), np=12, pencils=False),
    dict(name="io_pencils_512_np8", n=(512, 512, 512), np=8, io_pencils=True),
]


def fft3d_case_boxes(c):
    world = O.world_box(c["n"])
    oworld = world.r2c(c.get("r2c_dir", 0)) if c["kind"] == "r2c" else world
    return world, oworld, bricks(world, c["gin"]), bricks(oworld, c["gout"])


def plan_case_boxes(c):
    world = O.world_box(c["n"])
    r2c_dir = c.get("r2c_dir", -1)
    oworld = world.r2c(r2c_dir) if r2c_dir >= 0 else world
    if c.get("io_pencils"):
        g2 = ref_lib.make_procgrid(c["np"])
        gin, gout = (1, g2[0], g2[1]), (g2[0], g2[1], 1)
    else:
        gin = gout = tuple(ref_lib.proc_setup_min_surface(world, c["np"]))
    to_box = lambda nine: O.Box(nine[0:3], nine[3:6], nine[6:9])
    return [to_box(b) for b in ref_lib.split_world(world, gin)], [to_box(b) for b in ref_lib.split_world(oworld, gout)], gin, gout


def main():
    assert ref_lib.available() or ref_lib.build(), "oracle/_ref is not built"
    arrays, meta = {}, {}

    # the reference's input generator: minstd_rand(4242) -> U(0,1) (test/test_fft3d.h:19-38)
    arrays["make_data_64"] = ref_lib.make_data(64)

    for c in FFT3D_CASES:
        world, oworld, inboxes, outboxes = fft3d_case_boxes(c)
        x = ref_lib.make_data(world.count())
        if c["kind"] == "c2c":
            x = x + 1j * ref_lib.make_data(2 * world.count())[world.count():]
        inputs = [O.get_subbox(world, b, x) for b in inboxes]
        for scaling in ("none", "full", "symmetric"):
            outs, ws = ref_lib.fft3d(c["kind"], 1, inboxes, outboxes, inputs, backward=False, scaling=scaling, r2c_dir=c.get("r2c_dir", 0),
                                     use_reorder=c["reorder"], algorithm=c["alg"], use_pencils=c["pencils"])
            full = np.zeros(oworld.count(), dtype=outs[0].dtype)
            for b, o in zip(outboxes, outs):
                O.put_subbox(oworld, b, o, full)
            arrays["%s/forward_%s" % (c["name"], scaling)] = full
            if scaling == "none":
                backs, _ = ref_lib.fft3d(c["kind"], 1, inboxes, outboxes, outs, backward=True, scaling="none", r2c_dir=c.get("r2c_dir", 0),
                                         use_reorder=c["reorder"], algorithm=c["alg"], use_pencils=c["pencils"])
                fullb = np.zeros(world.count(), dtype=backs[0].dtype)
                for b, o in zip(inboxes, backs):
                    O.put_subbox(world, b, o, fullb)
                arrays["%s/backward_none" % c["name"]] = fullb
                meta[c["name"]] = dict(c, workspace=[int(v) for v in ws])
        arrays["%s/input" % c["name"]] = x

    for i, (kind, low, high, order, dim) in enumerate(EXEC1D_CASES):
        box = O.Box(low, high, order)
        x = ref_lib.make_data(2 * box.count())
        key = "exec1d_%02d_%s" % (i, kind)
        if kind == "c2c":
            z = x[:box.count()] + 1j * x[box.count():]
            arrays[key + "/input"] = z
            arrays[key + "/forward"] = ref_lib.exec1d_c2c(z, box, dim)
            arrays[key + "/backward"] = ref_lib.exec1d_c2c(z, box, dim, backward=True)
        elif kind == "r2c":
            r = x[:box.count()]
            arrays[key + "/input"] = r
            f = ref_lib.exec1d_r2c(r, box, dim)
            arrays[key + "/forward"] = f
            arrays[key + "/backward"] = ref_lib.exec1d_c2r(f, box, dim)
        else:
            r = x[:box.count()]
            arrays[key + "/input"] = r
            arrays[key + "/forward"] = ref_lib.exec1d_r2r(r, box, dim, kind)
            arrays[key + "/backward"] = ref_lib.exec1d_r2r(r, box, dim, kind, backward=True)
        meta[key] = dict(kind=kind, low=low, high=high, order=order, dim=dim)

    plans = {}
    for c in PLAN_CASES:
        inboxes, outboxes, gin, gout = plan_case_boxes(c)
        shapes, fdir, count = ref_lib.plan_operations(inboxes, outboxes, r2c_dir=c.get("r2c_dir", -1), use_reorder=c.get("reorder", False),
                                                      algorithm=c.get("alg", 0), use_pencils=c.get("pencils", True))
        plans[c["name"]] = dict(c, gin=list(gin), gout=list(gout), inboxes=[b.nine() for b in inboxes], outboxes=[b.nine() for b in outboxes],
                                shapes=shapes, fft_direction=fdir, index_count=count)

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **arrays)
    with open(os.path.join(HERE, "reference_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "reference_plans.json"), "w") as f:
        json.dump(plans, f, sort_keys=True)
    print("wrote %d arrays, %d plans" % (len(arrays), len(plans)))


if __name__ == "__main__":
    main()
