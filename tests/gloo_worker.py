"""
CPU multi-process check of the N > 1 host logic (gloo, no GPU): launched by tests/test_distributed_gloo.py under
torch.distributed.run.  Every rank
  1. builds a REAL plan through the C ABI (heffte_plan_create with a callback communicator whose allgather is a gloo
     all_gather): the plan-time exchange of boxes (reference include/heffte_geometry.h:707-718), the logic plan and
     the reshape send/receive lists are the product's own code;
  2. replays the stage sequence of the product (reshape s -> 1-D transform along fft_direction[s], reference
     src/heffte_compute_transform.cpp:15-258) on the host: messages are packed / unpacked with the numpy packers from
     the product's piece lists and travel through gloo send/recv; the 1-D transforms are the oracle's;
  3. compares its part of the result with the oracle's single-rank transform of the whole world.
This validates the routing (who sends what to whom, offsets, strides, permutation maps) that the CUDA path executes.
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import heffte_oracle as O  # noqa: E402
from tests.helpers import bricks, to_h  # noqa: E402


def gloo_allgather(mine):
    size = dist.get_world_size()
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
    out = [torch.empty_like(t) for _ in range(size)]
    dist.all_gather(out, t)
    return b"".join(bytes(o.numpy().tobytes()) for o in out)


def plan_of(piece):
    return dict(size=(piece["size0"], piece["size1"], piece["size2"]), line_stride=piece["line"], plane_stride=piece["plane"],
                buff_line_stride=piece["buff_line"], buff_plane_stride=piece["buff_plane"], map=(piece["map0"], piece["map1"], piece["map2"]))


def host_reshape(H, in_shape, out_shape, me, data, dtype):
    """the product's send/receive lists executed with numpy packers and gloo point-to-point messages"""
    ins = [to_h(b) for b in in_shape]
    outs = [to_h(b) for b in out_shape]
    sends = H.reshape_pieces(ins, outs, me, receive=False)
    recvs = H.reshape_pieces(ins, outs, me, receive=True)
    result = np.zeros(out_shape[me].count(), dtype=dtype)
    self_message = None
    requests, keep = [], []
    for s in sends:
        msg = O.direct_pack(plan_of(s), data, s["offset"])
        if s["peer"] == me:
            self_message = msg
        else:
            t = torch.from_numpy(np.ascontiguousarray(msg).view(np.uint8).copy())
            keep.append(t)
            requests.append(dist.isend(t, s["peer"]))
    for r in recvs:
        if r["peer"] == me:
            # a purely local re-ordering has no send list: the "message" is my whole box (reshape3d_transpose, heffte_reshape3d.h:472-484)
            msg = self_message if self_message is not None else data
        else:
            t = torch.empty(r["count"] * np.dtype(dtype).itemsize, dtype=torch.uint8)
            dist.recv(t, r["peer"])
            msg = t.numpy().view(dtype)
        permuted = (r["map0"], r["map1"], r["map2"]) != (0, 1, 2)
        if permuted:
            O.transpose_unpack(plan_of(r), msg, result, r["offset"])
        else:
            O.direct_unpack(plan_of(r), msg, result, r["offset"])
    for q in requests:
        q.wait()
    return result


def boxes_of(shape9):
    return [O.Box(b[0:3], b[3:6], b[6:9]) for b in shape9]


def same_shapes(a, b):
    return all(x.low == y.low and x.high == y.high and x.order == y.order for x, y in zip(a, b))


def run_case(H, rank, size, n, kind, gin, gout, reorder, pencils, r2c_dir=0):
    world = O.world_box(n)
    oworld = world.r2c(r2c_dir) if kind == "r2c" else world
    inboxes, outboxes = bricks(world, gin), bricks(oworld, gout)
    comm = H.comm_from_callbacks(rank, size, gloo_allgather)
    opts = H.plan_options(H.backend.b200, use_reorder=reorder, use_pencils=pencils)
    if kind == "r2c":
        fft = H.fft3d_r2c(H.backend.b200, to_h(inboxes[rank]), to_h(outboxes[rank]), r2c_dir, comm, opts)
    else:
        fft = H.fft3d(H.backend.b200, to_h(inboxes[rank]), to_h(outboxes[rank]), comm, opts)
    assert fft.size_inbox() == inboxes[rank].count() and fft.size_outbox() == outboxes[rank].count()
    expect_sizes = H.plan_sizes(1 if kind == "r2c" else 0, [to_h(b) for b in inboxes], [to_h(b) for b in outboxes], rank,
                                r2c_direction=r2c_dir if kind == "r2c" else -1, use_reorder=reorder, use_pencils=pencils)
    assert fft.size_workspace() == expect_sizes[2]

    shapes, fdir, _ = H.logic_plan([to_h(b) for b in inboxes], [to_h(b) for b in outboxes], r2c_direction=r2c_dir if kind == "r2c" else -1,
                                   use_reorder=reorder, use_pencils=pencils, rank=rank)
    rng = np.random.default_rng(99)
    x = rng.random(world.count())
    if kind == "c2c":
        x = x + 1j * rng.random(world.count())
    data = O.get_subbox(world, inboxes[rank], x)
    dtype = np.complex128 if kind == "c2c" else np.float64
    for s in range(4):
        ins, outs = boxes_of(shapes[s]), boxes_of(shapes[4 + s])
        if not same_shapes(ins, outs):
            data = host_reshape(H, ins, outs, rank, data, dtype)
        if s < 3:
            box = outs[rank]
            if box.count() > 0:
                if kind == "r2c" and s == 0:
                    data = O.exec1d_r2c(data, box, fdir[0])
                    dtype = np.complex128
                else:
                    data = O.exec1d_c2c(data, box, fdir[s])
            elif kind == "r2c" and s == 0:
                data, dtype = np.zeros(0, dtype=np.complex128), np.complex128
    ref = O.fft3d_forward(x, n, kind, r2c_dir=r2c_dir)
    expect = O.get_subbox(oworld, outboxes[rank], ref)
    err = O.rel_l2(data, expect) if expect.size else 0.0
    assert err < 1e-12, (n, kind, gin, gout, reorder, pencils, err)
    return err


def main():
    dist.init_process_group("gloo")
    rank, size = dist.get_rank(), dist.get_world_size()
    from heffte_b200 import heffte as H
    grids = {2: [((1, 1, 2), (2, 1, 1)), ((1, 2, 1), (1, 2, 1))], 4: [((1, 2, 2), (2, 2, 1)), ((4, 1, 1), (1, 1, 4))]}[size]
    count, worst = 0, 0.0
    for gin, gout in grids:
        for n in ((8, 8, 8), (9, 10, 12)):
            for reorder in (False, True):
                for pencils in (True, False):
                    worst = max(worst, run_case(H, rank, size, n, "c2c", gin, gout, reorder, pencils))
                    count += 1
                for r2c_dir in (0, 1, 2):
                    worst = max(worst, run_case(H, rank, size, n, "r2c", gin, gout, reorder, True, r2c_dir))
                    count += 1
    dist.barrier()
    if rank == 0:
        print("gloo_worker: ranks=%d cases=%d worst=%.2e ok" % (size, count, worst), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
