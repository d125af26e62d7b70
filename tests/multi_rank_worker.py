"""
Multi-rank GPU parity worker (one process per GPU, launched by torch.distributed.run from tests/test_gpu_multi.py or by
hand:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_rank_worker.py).

Mirrors the reference's distributed tests (test/test_fft3d.h:160-330: world array -> get_subbox per rank -> forward ->
compare with the matching sub-box of the single-rank result; test/test_fft3d_r2c.cpp; test/test_cos.cpp) with the numpy
oracle standing where the reference uses forward_fft<backend::stock>() on the whole world.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import heffte_oracle as O  # noqa: E402
from tests.helpers import TOL, bricks, to_h  # noqa: E402


def grids_for(nranks):
    """(in grid, out grid) pairs in the spirit of test/test_fft3d_np2/np4/np6/np8/np12.cpp"""
    table = {
        1: [((1, 1, 1), (1, 1, 1))],
        2: [((1, 1, 2), (1, 1, 2)), ((2, 1, 1), (1, 2, 1)), ((1, 2, 1), (2, 1, 1))],
        3: [((1, 3, 1), (3, 1, 1))],
        4: [((1, 2, 2), (1, 2, 2)), ((4, 1, 1), (1, 1, 4)), ((2, 2, 1), (1, 2, 2))],
        6: [((1, 2, 3), (3, 2, 1))],
        8: [((2, 2, 2), (2, 2, 2)), ((1, 2, 4), (2, 4, 1)), ((8, 1, 1), (1, 1, 8)), ((2, 4, 1), (1, 2, 4))],
    }
    return table.get(nranks, [((1, 1, nranks), (nranks, 1, 1))])


def configs(nranks, quick, subcomm=False):
    out = []
    sizes = [(16, 16, 16), (20, 21, 22)] if quick else [(16, 16, 16), (20, 21, 22), (64, 64, 64), (32, 48, 40)]
    for gi, (gin, gout) in enumerate(grids_for(nranks)):
        for n in sizes:
            for prec in (1, 0):
                for reorder in (False, True):
                    for pencils in (True, False):
                        for alg in (0, 3, 2, 1) if (not quick and gi == 0 and n == sizes[0]) else (0,):
                            out.append(dict(kind="c2c", n=n, prec=prec, reorder=reorder, pencils=pencils, alg=alg, gin=gin, gout=gout,
                                            order_out=(0, 1, 2)))
        # different order of the output boxes (test/test_fft3d.h:247-250 reordered io boxes)
        out.append(dict(kind="c2c", n=sizes[-1], prec=1, reorder=False, pencils=True, alg=0, gin=gin, gout=gout, order_out=(2, 0, 1)))
        for r2c_dir in (0, 1, 2):
            for prec in (1, 0):
                for reorder in (False, True):
                    out.append(dict(kind="r2c", n=sizes[-1], prec=prec, reorder=reorder, pencils=True, alg=0, gin=gin, gout=gout,
                                    order_out=(0, 1, 2), r2c_dir=r2c_dir))
        for kind in ("cos", "sin", "cos1"):
            out.append(dict(kind=kind, n=sizes[0], prec=1, reorder=True, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2)))
    # sub-communicators (test/test_subcomm.cpp): the intermediate stages live on the first ranks only
    if subcomm and nranks >= 4:
        gin, gout = grids_for(nranks)[0]
        for kind, sub in (("c2c", nranks // 2), ("r2c", 1), ("c2c", 3)):
            out.append(dict(kind=kind, n=sizes[0], prec=1, reorder=False, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2),
                            subranks=sub, r2c_dir=0))
    # power-of-two sizes that take the fast kernels on every stage
    n = (64, 64, 64) if quick else (128, 128, 128)
    gin, gout = grids_for(nranks)[0]
    for prec in (1, 0):
        out.append(dict(kind="c2c", n=n, prec=prec, reorder=False, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2)))
        out.append(dict(kind="r2c", n=n, prec=prec, reorder=False, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2), r2c_dir=0))
    return out


class TorchArrays:
    """device arrays of the GPU runs: torch CUDA tensors"""

    def __init__(self, torch):
        self.torch = torch

    def to_device(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def empty(self, count, dtype):
        return self.torch.empty(count, dtype=getattr(self.torch, np.dtype(dtype).name), device="cuda")

    def to_host(self, t):
        return t.cpu().numpy()


class HostArrays:
    """numpy arrays: the host entry point of the plan (used with the emulated library on the CPU)"""

    def to_device(self, a):
        return np.ascontiguousarray(a).copy()

    def empty(self, count, dtype):
        return np.zeros(count, dtype=dtype)

    def to_host(self, t):
        return t


_memo, _memo_lock = {}, __import__("threading").Lock()


def _shared(key, compute):
    """the oracle answers of a configuration, computed once per process: ranks that are threads of one process (the emulated and
    the single-GPU thread-rank tests) would otherwise repeat the same numpy transforms under the GIL, one after the other"""
    with _memo_lock:
        slot = _memo.get(key)
        if slot is None:
            slot = {"lock": __import__("threading").Lock(), "value": None}
            _memo[key] = slot
            if len(_memo) > 64:
                for old in list(_memo)[:32]:
                    if old != key:
                        _memo.pop(old, None)
    with slot["lock"]:
        if slot["value"] is None:
            slot["value"] = compute()
        return slot["value"]


def run_config(hf, torch, comm, rank, c, batch=1, stream=None, expect_peer=None, arrays=None):
    arrays = arrays or TorchArrays(torch)
    n, kind, prec = c["n"], c["kind"], c["prec"]
    world = O.world_box(n)
    r2c_dir = c.get("r2c_dir", 0)
    oworld = world.r2c(r2c_dir) if kind == "r2c" else world
    inboxes = bricks(world, c["gin"])
    outboxes = bricks(oworld, c["gout"], c["order_out"])
    inbox, outbox = inboxes[rank], outboxes[rank]
    complex_in = kind == "c2c"
    rng = np.random.default_rng(1234)
    x = rng.random(world.count())
    if complex_in:
        x = x + 1j * rng.random(world.count())
    rdt, cdt = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    x = x.astype(cdt if complex_in else rdt)
    tag = {"c2c": hf.backend.b200, "r2c": hf.backend.b200, "cos": hf.backend.b200_cos, "sin": hf.backend.b200_sin, "cos1": hf.backend.b200_cos1}[kind]
    opts = hf.plan_options(tag, use_reorder=c["reorder"], algorithm=c["alg"], use_pencils=c["pencils"])
    if c.get("subranks"):
        opts.use_subcomm(c["subranks"])      # test/test_subcomm.cpp: intermediate stages on fewer ranks
    if kind == "r2c":
        fft = hf.fft3d_r2c(tag, to_h(inbox), to_h(outbox), r2c_dir, comm, opts, stream=stream)
    else:
        fft = hf.fft3d(tag, to_h(inbox), to_h(outbox), comm, opts, stream=stream)
    assert fft.size_inbox() == inbox.count() and fft.size_outbox() == outbox.count()
    local = O.get_subbox(world, inbox, x)
    out_dtype = cdt if kind in ("c2c", "r2c") else rdt
    tol = TOL[prec] * (4 if kind in ("cos", "sin", "cos1") else 1)
    worst = 0.0
    problems = []   # reported at the end: every rank must issue the same sequence of collective calls whatever the numbers are
    for scaling, sname in ((1, "full"), (0, "none")):
        ref = _shared(("fwd", kind, n, prec, r2c_dir, sname), lambda: O.fft3d_forward(x, n, kind, r2c_dir=r2c_dir, scaling=sname))
        dx = arrays.to_device(np.tile(local, batch))
        dy = arrays.empty(batch * outbox.count(), out_dtype)
        fft.forward(dx, dy, scaling, batch=batch)
        expect = O.get_subbox(oworld, outbox, ref)
        got = arrays.to_host(dy)
        for b in range(batch):
            seg = got[b * outbox.count():(b + 1) * outbox.count()]
            err = O.rel_l2(seg, expect) if expect.size else 0.0
            worst = max(worst, err)
            if not err <= tol:
                problems.append("forward(%s): rel l2 %.3e > %.1e" % (sname, err, tol))
        dz = arrays.empty(batch * inbox.count(), x.dtype)
        fft.backward(dy, dz, scaling, batch=batch)
        refb = _shared(("bwd", kind, n, prec, r2c_dir, sname), lambda: O.fft3d_backward(ref, n, kind, r2c_dir=r2c_dir, scaling=sname))
        expect_b = O.get_subbox(world, inbox, refb)
        got = arrays.to_host(dz)
        for b in range(batch):
            seg = got[b * inbox.count():(b + 1) * inbox.count()]
            err = O.rel_l2(seg, expect_b) if expect_b.size else 0.0
            worst = max(worst, err)
            if not err <= 2 * tol:
                problems.append("backward(%s): rel l2 %.3e > %.1e" % (sname, err, 2 * tol))
    if expect_peer is not None and comm.size() > 1 and fft.uses_peer_memory(prec) != expect_peer:
        problems.append("peer-memory mode is %s" % fft.uses_peer_memory(prec))
    # in-place with a caller workspace (c2c and r2r only), the way speed3d drives the plan
    # (a global condition: the transforms are collective, every rank must take the same decision)
    if kind != "r2c" and all(a.count() == b.count() for a, b in zip(inboxes, outboxes)):
        work = arrays.empty(fft.size_workspace(), x.dtype)
        d = arrays.to_device(local)
        fft.forward_buffered(d, d, work, 1)
        fft.backward_buffered(d, d, work, 0)
        err = O.rel_l2(arrays.to_host(d), local) if local.size else 0.0
        if not err <= 2 * tol:
            problems.append("in-place round trip: %.3e" % err)
    # caller arrays registered with the plan: the other GPUs store their part of the result straight into them (peer-memory
    # mode; elsewhere the registration is refused on every rank alike and the transforms run as before)
    if comm.size() > 1 and batch == 1:
        dx = arrays.to_device(local)
        dy = arrays.empty(outbox.count(), out_dtype)
        dz = arrays.empty(inbox.count(), x.dtype)
        registered = [fft.register_buffer(dy), fft.register_buffer(dz)]
        if isinstance(arrays, HostArrays):
            # numpy arrays take the host entry point (staged through the plan's own buffers); with the emulated library host
            # memory IS device memory, so the device entry point can be driven directly and the registered arrays are reached
            import ctypes
            from heffte_b200 import _lib as L

            def run_device(direction, a, b, scaling):
                rc = L.load().heffte_execute(fft.plan, prec, direction, 1, ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(b.ctypes.data), None, scaling)
                assert rc == 0, L.last_error()
            forward_, backward_ = (lambda a, b, s: run_device(0, a, b, s)), (lambda a, b, s: run_device(1, a, b, s))
        else:
            forward_, backward_ = fft.forward, fft.backward
        # (a registration is also refused when the last reshape of a direction moves nothing: there is nothing to store remotely)
        if any(registered) and not fft.uses_peer_memory(prec):
            problems.append("register_buffer answered %s in exchange mode" % registered)
        forward_(dx, dy, 0)
        err = O.rel_l2(arrays.to_host(dy), expect) if expect.size else 0.0
        backward_(dy, dz, 0)
        errb = O.rel_l2(arrays.to_host(dz), expect_b) if expect_b.size else 0.0
        worst = max(worst, err, errb)
        if not (err <= tol and errb <= 2 * tol):
            problems.append("registered output arrays: forward %.3e backward %.3e" % (err, errb))
        if kind != "r2c" and all(a.count() == b.count() for a, b in zip(inboxes, outboxes)):
            forward_(dz, dz, 1)                     # in place into a registered array
            backward_(dz, dz, 0)
            err = O.rel_l2(arrays.to_host(dz), expect_b) if expect_b.size else 0.0
            if not err <= 4 * tol:
                problems.append("registered array, in-place round trip: %.3e" % err)
    # fused spectral operator (reference benchmarks/convolution.cpp:86-97): forward(scale full), pointwise product, backward, as
    # ONE plan call -- the spectrum times itself, then times a caller array laid out over convolve_box()
    if kind == "c2c" and batch == 1 and all(a.count() == b.count() for a, b in zip(inboxes, outboxes)):
        spectrum = _shared(("fwd", kind, n, prec, r2c_dir, "full"), lambda: O.fft3d_forward(x, n, "c2c", scaling="full"))
        for use_multiplier in (False, True):
            lo, hi, order = fft.convolve_box()
            cbox = O.Box(lo, hi, order)
            mult_world = None
            dm = None
            if use_multiplier:
                mrng = np.random.default_rng(77)
                mult_world = (mrng.random(world.count()) + 1j * mrng.random(world.count())).astype(cdt)
                dm = arrays.to_device(O.get_subbox(world, cbox, mult_world))
            product = spectrum * (mult_world if use_multiplier else spectrum)
            conv_world = _shared(("conv", n, prec, use_multiplier), lambda: O.fft3d_backward(product, n, "c2c", scaling="none"))
            expect_c = O.get_subbox(world, inbox, conv_world)
            d = arrays.to_device(local)
            dout = arrays.empty(inbox.count(), x.dtype)
            try:
                fft.convolve(d, dout, dm, 1)
            except Exception as e:  # noqa: BLE001  (the exchange path offers the self-product only)
                if use_multiplier and "multiplier" in str(e):
                    continue
                raise
            err = O.rel_l2(arrays.to_host(dout), expect_c) if expect_c.size else 0.0
            worst = max(worst, err)
            if not err <= 4 * tol:
                problems.append("convolve(%s): rel l2 %.3e" % ("multiplier" if use_multiplier else "self", err))
    assert not problems, "rank %d %s: %s" % (rank, c, "; ".join(problems))
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--subcomm", action="store_true", help="also run the sub-communicator plans (test/test_subcomm.cpp)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import heffte_b200 as hf
    rank, size = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = hf.comm_from_torch()
    else:
        comm = hf.comm_self()
    done, worst = 0, 0.0
    failed = None
    gin, gout = grids_for(size)[0]
    todo = [(c, 1) for c in configs(size, args.quick, subcomm=args.subcomm)]
    # batched transforms across ranks (test/test_fft3d.h:505-572)
    todo.append((dict(kind="c2c", n=(16, 18, 20), prec=1, reorder=False, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2)), 3))
    # lines of 128 / 256 points along the two fast axes: the local transform in front of a fused stage runs inside its persistent
    # kernel (fft_pair_kernel), single and batched
    todo.append((dict(kind="c2c", n=(128, 128, 32), prec=0, reorder=False, pencils=False, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2)), 2))
    todo.append((dict(kind="c2c", n=(256, 256, 16), prec=1, reorder=False, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2)), 1))
    flag = torch.zeros(1, device="cuda", dtype=torch.int32)
    for c, batch in todo:
        try:
            worst = max(worst, run_config(hf, torch, comm, rank, c, batch))
            done += 1
        except AssertionError as e:
            failed = str(e)
        except Exception as e:  # noqa: BLE001
            failed = repr(e)
        # every rank learns about a failure before the next collective plan creation (no hangs)
        flag.fill_(1 if failed else 0)
        if size > 1:
            dist.all_reduce(flag)
        if int(flag.item()):
            break
    if failed:
        print("rank %d FAILED after %d configs: %s" % (rank, done, failed), flush=True)
    if rank == 0:
        print("multi_rank_worker: ranks=%d configs=%d worst_rel_l2=%.3e failures=%d" % (size, done, worst, int(flag.item())), flush=True)
    if size > 1:
        dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
