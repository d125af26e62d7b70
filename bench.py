#!/usr/bin/env python
"""
bench.py -- the measurement harness: heFFTe's speed3d metric for the b200 backend.

Metric (reference benchmarks/speed3d.h:168-195, 224-230): one *step* = forward(scale::full) + backward on the same
buffers; t = time / (2 * steps); GFlop/s = 5 * N * log2(N) * 1e-9 / t with N = nx*ny*nz, whole job (all ranks).
Default workload: speed3d_c2c double 512^3 in place, bricks on the proc_setup_min_surface grid (1 GPU: the whole box).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # b200 arm (N > 1: launched by torchrun)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # the reference's own CPU path (oracle/_ref)

Prints ONE JSON line on rank 0.  `value` = device-resident throughput, `e2e` = same metric through the host-buffer
entry point (heffte_execute_host: H2D + transform + D2H per call), `roofline` = dominant FFT kernel against the
measured HBM peak (N = 1) or the dominant fused stage against the NVLink peer-copy rate (N > 1), `cpu_baseline` = the
unmodified reference (stock backend, threads-as-ranks MPI stand-in) timed on this box's host cores, `parity_rel_l2` =
forward(scale::full) of a hashed world array on the GPUs against the SAME transform by the compiled reference
(oracle/_ref as the checker, outside every timed region; all ranks' sub-boxes), `secondary` = the other BASELINE.json
configurations (r2c, DCT, convolution, fp32 256^3 with an L2 flush, fp32 1024^3 on 8 GPUs) with value, roofline
fraction and parity each.
"""
import argparse
import ctypes
import json
import math
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NVLINK_PEAK = 770.0   # GB/s per direction per GPU: measured peer copy on this pool (B200_PROFILING.md; tools/ipc_probe.py saw 760)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, nargs=3, default=[512, 512, 512])
    ap.add_argument("--precision", default="double", choices=["float", "double"])
    ap.add_argument("--kind", default="c2c", choices=["c2c", "r2c", "r2r", "conv"],
                    help="r2r = DCT-II / DCT-III (speed3d_r2r ... cos); conv = benchmarks/convolution.cpp: forward(scale full), x *= x, backward, c2c in place")
    ap.add_argument("--reorder", action="store_true")
    ap.add_argument("--slabs", action="store_true", help="speed3d -slabs: the slab decomposition is executed (default: the planner picks)")
    ap.add_argument("--pencils", action="store_true", help="speed3d -pencils: the pencil decomposition is executed")
    ap.add_argument("--io-pencils", action="store_true", help="pencil-shaped in/out boxes (speed3d -io_pencils)")
    ap.add_argument("--batch", type=int, default=1, help="speed3d -batch: transforms per call")
    ap.add_argument("--flush-l2", action="store_true", help="write a 512 MB buffer between timed steps (forced when the working set fits the L2)")
    ap.add_argument("--l2-slab-mb", type=float, default=None,
                    help="experimental: run pairs of local transforms slab by slab through the L2 cache (sets HEFFTE_B200_L2_SLAB_MB)")
    ap.add_argument("--unfused-conv", action="store_true", help="--kind conv as three caller-side calls (forward, multiply, backward) instead of plan.convolve()")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the comparison with the compiled reference (oracle/_ref)")
    ap.add_argument("--no-secondary", action="store_true", help="primary workload only")
    ap.add_argument("--no-register", action="store_true", help="several GPUs: do not register the benchmark's arrays with the plan (the results then pass through the plan's buffers)")
    ap.add_argument("--cpu-size", type=int, nargs=3, default=None, help="size of the CPU sample (default: the workload itself)")
    return ap.parse_args()


def gflops(n, seconds_per_transform, batch=1):
    N = float(n[0]) * n[1] * n[2]
    return 5.0 * batch * N * math.log2(N) * 1e-9 / seconds_per_transform


def workload_name(kind, precision, size):
    """identical in both arms (the driver compares the strings); everything else about a run sits under other config keys"""
    return "speed3d_%s %s %dx%dx%d" % (kind, precision, size[0], size[1], size[2])


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML (same counters as the nvidia-smi clocks line) every 5 ms."""

    def __init__(self, device_index):
        self.device_index = device_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.running = False
        self.thread = None
        self.error = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = self.device_index
            if visible:
                try:
                    index = int(visible.split(",")[self.device_index])
                except Exception:
                    pass
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.running = True
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception as e:
            self.error = repr(e)

    def _loop(self):
        n = self.nvml
        masks = {}
        for name, attr in [("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                           ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")]:
            alt = attr.replace("ClocksEventReason", "ClocksThrottleReason")
            masks[name] = getattr(n, attr, getattr(n, alt, 0))
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
        while self.running:
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    bits = get_reasons(self.handle)
                    for name, mask in masks.items():
                        if mask and (bits & mask):
                            self.reasons.add(name)
            except Exception as e:
                self.error = repr(e)
                break
            time.sleep(0.005)

    def stop(self):
        self.running = False
        if self.thread is not None:
            self.thread.join(timeout=2)
        sm = sorted(self.samples)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm)}
        if self.error:
            out["error"] = self.error
        return out


# ----------------------------------------------------------------------------------------------------------------
# the reference arm / cpu baseline: oracle/_ref speed3d binaries (the unmodified reference, stock backend)
# ----------------------------------------------------------------------------------------------------------------
def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_speed3d(kind, precision, size, nruns, options=()):
    """Runs the reference's own speed3d_<kind> (stock backend) on thread-ranks; returns dict or None."""
    from oracle import ref_lib
    binary = ref_lib.binary("speed3d_" + ("c2c" if kind == "conv" else kind))
    if binary is None or not os.path.exists(binary):
        return None
    cores = usable_cores()
    ranks = 1
    while ranks * 2 <= min(cores, 64):
        ranks *= 2
    env = dict(os.environ, SHIM_NP=str(ranks))
    cmd = [binary, "stock-cos" if kind == "r2r" else "stock", precision, str(size[0]), str(size[1]), str(size[2]), "-n%d" % nruns] + list(options)
    t0 = time.time()
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=3000)
    wall = time.time() - t0
    m = re.search(r"Performance:\s+([0-9.eE+-]+)\s+GFlops/s", out.stdout)
    t = re.search(r"Time per run:\s+([0-9.eE+-]+)", out.stdout)
    if out.returncode != 0 or m is None:
        return {"error": (out.stdout + out.stderr)[-400:]}
    return {"gflops": float(m.group(1)), "seconds_per_transform": float(t.group(1)) if t else None, "ranks": ranks,
            "cores": cores, "wall_s": wall, "isa": os.path.basename(os.path.dirname(binary)), "nruns": nruns,
            "sample": "speed3d_%s stock %s %dx%dx%d -n%d on %d thread-ranks (reference compiled in place, %s)" % (
                kind, precision, size[0], size[1], size[2], nruns, ranks, os.path.basename(os.path.dirname(binary)))}


def reference_arm(args):
    """
    The reference's own benchmark binary, unmodified.  Its timed loop (benchmarks/speed3d.h:163-191) is ONE untimed
    forward+backward pair followed by -n<steps> timed pairs; the line reports the steps and warm-up that really ran.
    A first -n1 run sizes the sample: the timed run is cut to what fits in about two minutes.
    """
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size = args.cpu_size or args.size
    kind = "c2c" if args.kind == "conv" else args.kind
    probe = run_reference_speed3d(kind, args.precision, size, 1)
    if probe is None or "error" in probe:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref speed3d binary missing or failed: %s" % (probe or {}).get("error", "not built")}))
        return
    per_step = max(2.0 * probe["seconds_per_transform"], 1e-6)
    steps = max(1, min(args.steps, int(120.0 / per_step)))
    best = probe if steps == 1 else run_reference_speed3d(kind, args.precision, size, steps)
    if best is None or "error" in best:
        best, steps = probe, 1
    line = {
        "impl": "reference", "metric": "speed3d_%s GFlop/s (5*N*log2(N)/t)" % args.kind, "value": best["gflops"], "unit": "GFlop/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": 2e3 * best["seconds_per_transform"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if args.precision == "double" else "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.kind, args.precision, size),
                   "backend": "stock (FFTW and MPI are absent from the image)", "ranks": best["ranks"],
                   "requested": {"steps": args.steps, "warmup": args.warmup},
                   "note": "the reference binary runs one untimed pair, then -n<steps> timed pairs; a -n1 probe run precedes it"},
        "cpu_baseline": {"value": best["gflops"], "unit": "GFlop/s", "cores": best["ranks"], "kind": "reference", "sample": best["sample"]},
        "e2e": {"value": best["gflops"], "unit": "GFlop/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# parity against the compiled reference (the checker; never inside a timed region)
# ----------------------------------------------------------------------------------------------------------------
HASH_K = 0x9E3779B97F4A7C15
HASH_K_SIGNED = HASH_K - (1 << 64)


def hashed_world_numpy(count, salt):
    """u[i] = 24 bits of the product (i + salt) * K (mod 2^64), in [0, 1): exactly representable in fp32 and fp64"""
    import numpy as np
    idx = np.arange(count, dtype=np.uint64) + np.uint64(salt)
    return (((idx * np.uint64(HASH_K)) >> np.uint64(20)) & np.uint64(0xFFFFFF)).astype(np.float64) * (1.0 / 16777216.0)


def hashed_box_torch(torch, box, n, salt, dtype):
    """the same numbers for the elements of `box` (order (0,1,2)) of the world n, computed where the tensors live"""
    lo, hi = [int(v) for v in box.low], [int(v) for v in box.high]
    if any(h < l for l, h in zip(lo, hi)):
        return torch.zeros(0, dtype=dtype, device="cuda")
    i0 = torch.arange(lo[0], hi[0] + 1, dtype=torch.int64, device="cuda")
    i1 = torch.arange(lo[1], hi[1] + 1, dtype=torch.int64, device="cuda")
    i2 = torch.arange(lo[2], hi[2] + 1, dtype=torch.int64, device="cuda")
    idx = (i2[:, None, None] * n[1] + i1[None, :, None]) * n[0] + i0[None, None, :] + salt
    h = ((idx * HASH_K_SIGNED) >> 20) & 0xFFFFFF       # int64 wraps like uint64; bits 20..43 do not see the sign extension
    return (h.to(torch.float64) * (1.0 / 16777216.0)).to(dtype).reshape(-1)


def reference_world(kind, n, x_world, salt_unused=0):
    """forward(scale::full) [conv: then x*x, backward] of the world array by oracle/_ref on thread-ranks; None when unavailable"""
    try:
        from oracle import ref_lib
        from tests.helpers import bricks, host_ranks
        from oracle import heffte_oracle as O
        if not ref_lib.available():
            return None, "oracle/_ref not built on this host"
        import numpy as np
        ranks = max(1, min(host_ranks(), n[2]))
        world = O.world_box(n)
        rkind = {"c2c": "c2c", "conv": "c2c", "r2c": "r2c", "r2r": "cos"}[kind]
        out_world = world.r2c(0) if kind == "r2c" else world
        inb, outb = bricks(world, (1, 1, ranks)), bricks(out_world, (1, 1, ranks))
        plane = n[0] * n[1]
        inputs = [x_world[b.low[2] * plane:(b.high[2] + 1) * plane] for b in inb]
        outs, _ = ref_lib.fft3d(rkind, 1, inb, outb, inputs, scaling="full", r2c_dir=0)
        if kind == "conv":
            outs = [o * o for o in outs]
            outs, _ = ref_lib.fft3d("c2c", 1, inb, outb, outs, backward=True, scaling="none")
        return np.concatenate(outs), "heffte::fft3d<stock> fp64 (oracle/_ref) on %d thread-ranks" % ranks
    except Exception as e:  # noqa: BLE001  (the checker must never take the bench down)
        return None, "reference checker failed: %r" % (e,)


# ----------------------------------------------------------------------------------------------------------------
# the b200 arm
# ----------------------------------------------------------------------------------------------------------------
class Context:
    pass


def make_context(args):
    import torch
    import heffte_b200 as hf
    from heffte_b200 import _lib, build
    ctx = Context()
    ctx.torch, ctx.hf = torch, hf
    ctx.world_size = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the library is built in-tree ahead of time (python -m heffte_b200.build / __graft_entry__.build()); build here only if it is
    # missing, and never from several ranks at once
    if not os.path.exists(build.library_path()):
        if ctx.world_size > 1:
            raise SystemExit("bench.py: %s is missing; run `python -m heffte_b200.build` before a multi-rank launch" % build.library_path())
        build.build_library()
    ctx.lib = _lib.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 backend has no CPU fallback")
    torch.cuda.set_device(ctx.local_rank)
    ctx.distributed = ctx.world_size > 1
    if ctx.distributed:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local_rank))
        ctx.dist = dist
        ctx.comm = hf.comm_from_torch()
    else:
        ctx.dist = None
        ctx.comm = hf.comm_self()
    ctx.flush = None
    return ctx


def barrier(ctx):
    if ctx.distributed:
        ctx.dist.barrier()
    ctx.torch.cuda.synchronize()


def all_max(ctx, values):
    if not ctx.distributed:
        return list(values)
    t = ctx.torch.tensor(list(values), dtype=ctx.torch.float64, device="cuda")
    ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
    return t.tolist()


def all_sum(ctx, values):
    if not ctx.distributed:
        return list(values)
    t = ctx.torch.tensor(list(values), dtype=ctx.torch.float64, device="cuda")
    ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.SUM)
    return t.tolist()


def flush_l2(ctx):
    """overwrite a 512 MB buffer: four times the 126 MB L2"""
    torch = ctx.torch
    if ctx.flush is None:
        ctx.flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    ctx.flush.fill_(1.0)


def run_workload(ctx, args, kind, size, precision, steps, warmup, *, primary, reorder=False, slabs=False, pencils=False, io_pencils=False, batch=1,
                 force_flush=False, parity=True, e2e_wanted=True, cpu_wanted=True):
    torch, hf, lib, dist = ctx.torch, ctx.hf, ctx.lib, ctx.dist
    rank, world_size, distributed = ctx.rank, ctx.world_size, ctx.distributed
    n = tuple(size)
    prec = 0 if precision == "float" else 1
    rdtype = torch.float32 if prec == 0 else torch.float64
    cdtype = torch.complex64 if prec == 0 else torch.complex128
    world = hf.box3d((0, 0, 0), (n[0] - 1, n[1] - 1, n[2] - 1))

    if io_pencils:
        g2 = hf.heffte.make_procgrid(world_size)
        in_grid, out_grid = [1, g2[0], g2[1]], [g2[0], g2[1], 1]
    else:
        in_grid = out_grid = hf.heffte.proc_setup_min_surface(world, world_size)
    r2c, r2r, conv = kind == "r2c", kind == "r2r", kind == "conv"
    inbox = hf.heffte.split_world(world, in_grid)[rank]
    cworld = hf.box3d((0, 0, 0), (n[0] // 2, n[1] - 1, n[2] - 1))
    outbox = hf.heffte.split_world(cworld if r2c else world, out_grid)[rank]

    tag = hf.backend.b200_cos if r2r else hf.backend.b200
    decomposition = False if slabs else (True if pencils else None)       # None: the planner picks the cheaper one
    options = hf.plan_options(tag, use_reorder=reorder, use_pencils=decomposition)
    fft = hf.fft3d_r2c(tag, inbox, outbox, 0, ctx.comm, options) if r2c else hf.fft3d(tag, inbox, outbox, ctx.comm, options)

    # the plan that runs (pure host planning, csrc/plan_logic.h): process grids of input, the three transform stages, output
    try:
        world_boxes_in = hf.heffte.split_world(world, in_grid)
        world_boxes_out = hf.heffte.split_world(cworld if r2c else world, out_grid)
        shapes, _, swaps = hf.heffte.execution_plan(world_boxes_in, world_boxes_out, r2c_direction=0 if r2c else -1,
                                                    use_reorder=bool(reorder or r2r), use_pencils=decomposition)

        def grid_of(boxes):
            return "x".join(str(len({(b[d], b[3 + d]) for b in boxes if all(b[3 + k] >= b[k] for k in range(3))})) for d in range(3))
        executed = {"grids": " -> ".join(grid_of(shapes[i]) for i in (0, 4, 5, 6, 7)), "refinements_applied": swaps}
    except Exception as e:  # noqa: BLE001  (reporting only)
        executed = {"error": repr(e)}

    nin, nout = fft.size_inbox(), fft.size_outbox()
    work = torch.empty(fft.size_workspace() * batch, dtype=rdtype if r2r else cdtype, device="cuda")
    register = distributed and batch == 1 and not args.no_register

    # ---- parity: forward(scale::full) of a hashed world array against the compiled reference, every rank's sub-box --------
    parity_info = None
    count = n[0] * n[1] * n[2]
    if parity and not args.no_parity:
        too_big = count > 512 ** 3
        real_in = r2c or r2r
        xin = hashed_box_torch(torch, inbox, n, 0, rdtype)
        if not real_in:
            xin = torch.complex(xin, hashed_box_torch(torch, inbox, n, count, rdtype))
        if conv:
            # the fused spectral operator: forward(scale full), spectrum times itself, backward, one plan-level call
            yout = torch.empty(nin, dtype=cdtype, device="cuda")
            checked_registered = register and fft.register_buffer(yout)
            fft.convolve(xin, yout, None, hf.scale.full)
        else:
            yout = torch.empty(nout, dtype=rdtype if r2r else cdtype, device="cuda")
            checked_registered = register and fft.register_buffer(yout)      # the timed loop writes registered arrays: so does the check
            fft.forward(xin, yout, hf.scale.full)
        torch.cuda.synchronize()
        if checked_registered:
            fft.unregister_buffer(yout)
        expect_dev, checker = None, None
        if too_big:
            checker = "skipped: the reference needs more than 64 GB of host memory for this size"
        else:
            flag = [0.0]
            if rank == 0:
                import numpy as np
                xw = hashed_world_numpy(count, 0)
                if not real_in:
                    xw = xw + 1j * hashed_world_numpy(count, count)
                expect, checker = reference_world(kind, n, xw)
                del xw
                if expect is not None:
                    expect_dev = torch.from_numpy(np.ascontiguousarray(expect)).cuda()
                    flag = [1.0]
                    del expect
            flag = all_max(ctx, flag)
            if flag[0] > 0:
                full = cworld if r2c else world
                shape = (int(full.size[2]), int(full.size[1]), int(full.size[0]))
                if expect_dev is None:
                    expect_dev = torch.empty(shape[0] * shape[1] * shape[2], dtype=torch.float64 if r2r else torch.complex128, device="cuda")
                if distributed:
                    dist.broadcast(torch.view_as_real(expect_dev) if expect_dev.is_complex() else expect_dev, src=0)
                box = inbox if conv else outbox
                lo, hi = [int(v) for v in box.low], [int(v) for v in box.high]
                mine = expect_dev.reshape(shape)[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1].reshape(-1)
                got = yout.to(mine.dtype)
                num = float((got - mine).abs().square().sum().item())
                den = float(mine.abs().square().sum().item())
                num, den = all_sum(ctx, [num, den])
                rel = math.sqrt(num / den) if den > 0 else float("nan")
                tol = (1e-5 if prec == 0 else 1e-12) * (4 if r2r else 1)
                parity_info = {"parity_rel_l2": rel, "parity_ok": bool(rel <= tol), "tolerance": tol,
                               "checker": checker, "compared": "forward(scale full)%s of a hashed world array, all %d ranks' sub-boxes" % (
                                   " * itself, backward" if conv else "", world_size)}
                del mine, got
            else:
                parity_info = {"parity_rel_l2": None, "parity_ok": None, "checker": None}
                if rank == 0:
                    parity_info["checker"] = checker
            del expect_dev
        if parity_info is None:
            parity_info = {"parity_rel_l2": None, "parity_ok": None, "checker": checker}
        del xin, yout
        torch.cuda.empty_cache()

    # ---- the speed3d buffers -----------------------------------------------------------------------------------------------
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4242 + rank)
    if r2c:
        data_in = torch.rand(nin * batch, dtype=rdtype, device="cuda", generator=gen)
        data_out = torch.empty(nout * batch, dtype=cdtype, device="cuda")
    elif r2r:
        data_in = torch.rand(max(nin, nout) * batch, dtype=rdtype, device="cuda", generator=gen)
        data_out = data_in
    else:
        # speed3d: complex data with zero imaginary part, transformed in place
        data_in = torch.complex(torch.rand(max(nin, nout) * batch, dtype=rdtype, device="cuda", generator=gen),
                                torch.zeros(max(nin, nout) * batch, dtype=rdtype, device="cuda"))
        data_out = data_in
    reference_copy = data_in.clone()
    # several GPUs: the arrays of the benchmark are registered with the plan (once, like a buffer registered with a communication
    # library): the other GPUs store their part of every result straight into them
    registered = False
    if register:
        registered = bool(fft.register_buffer(data_in))
        if data_out is not data_in:
            registered = bool(fft.register_buffer(data_out)) or registered

    fused_conv = conv and not args.unfused_conv

    def step():
        if fused_conv:
            fft.convolve_buffered(data_in, data_in, work, None, hf.scale.full)   # forward, x *= x, backward: one plan-level call
            return
        fft.forward_buffered(data_in, data_out, work, hf.scale.full, batch)
        if conv:
            data_out.mul_(data_out)      # the caller's pointwise product in spectral space (benchmarks/convolution.cpp:89-94)
        fft.backward_buffered(data_out, data_in, work, hf.scale.none, batch)

    for _ in range(max(warmup, 3)):
        step()
    barrier(ctx)
    # accuracy of the round trip (benchmarks/speed3d.h:213-220)
    err = float((data_in - reference_copy).abs().max().item()) if not conv else float("nan")
    data_in.copy_(reference_copy)

    working_set_mb = max(nin, nout) * batch * ((4 if prec == 0 else 8) * (1 if r2r else 2)) / 1e6
    flush = bool(force_flush or args.flush_l2 or working_set_mb <= 2 * 126)
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and primary:
        sampler.start()
    launches0 = lib.b200_launch_count()
    if flush:
        # the working set is near the 126 MB L2: flush between the steps, events around every step, flush time not counted
        events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier(ctx)
        for a, b in events:
            flush_l2(ctx)
            a.record()
            step()
            b.record()
        barrier(ctx)
        elapsed_ms = sum(a.elapsed_time(b) for a, b in events)
    else:
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(ctx)
        start.record()
        for _ in range(steps):
            step()
        stop.record()
        barrier(ctx)
        elapsed_ms = start.elapsed_time(stop)
    launches = lib.b200_launch_count() - launches0
    clocks = sampler.stop() if (rank == 0 and primary) else None
    elapsed_ms, err = all_max(ctx, [elapsed_ms, err])
    ms_per_step = elapsed_ms / steps
    sec_per_transform = ms_per_step * 1e-3 / 2.0
    value = gflops(n, sec_per_transform, batch)

    # ---- end to end: host buffers in, host buffers out, through the public plan API ---------------------------------
    e2e = None
    if e2e_wanted and not args.no_e2e and not conv and batch == 1:
        host_in = torch.empty(nin if r2c else max(nin, nout), dtype=rdtype if (r2c or r2r) else cdtype).pin_memory()
        host_mid = torch.empty(nout if r2c else max(nin, nout), dtype=rdtype if r2r else cdtype).pin_memory()
        host_in.copy_(reference_copy.cpu())
        np_in, np_mid = host_in.numpy(), host_mid.numpy()
        np_back = torch.empty_like(host_in).pin_memory().numpy()
        e2e_steps = max(2, min(steps, 5))

        def e2e_step():
            fft.forward(np_in, np_mid, hf.scale.full)      # H2D(in) + forward + D2H(out)
            fft.backward(np_mid, np_back, hf.scale.none)   # H2D(out) + backward + D2H(in)

        e2e_step()
        barrier(ctx)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier(ctx)
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        e2e_s = all_max(ctx, [e2e_s])[0]
        in_bytes = host_in.numel() * host_in.element_size()
        mid_bytes = host_mid.numel() * host_mid.element_size()
        e2e = {"value": gflops(n, e2e_s / 2.0), "unit": "GFlop/s", "h2d_bytes_per_step": in_bytes + mid_bytes,
               "d2h_bytes_per_step": in_bytes + mid_bytes, "steps": e2e_steps,
               "pcie_GB/s_per_gpu": 2.0 * (in_bytes + mid_bytes) / e2e_s * 1e-9,
               "note": "per rank bytes; step = forward(host in -> host out) + backward(host out -> host in), pinned buffers"}
        del host_in, host_mid, np_in, np_mid, np_back

    # ---- multi-GPU: device time of every stage of one forward and one backward transform (CUDA events on the plan's stream) ----
    multi = None
    if distributed and batch == 1:
        peer_mode = bool(fft.uses_peer_memory(prec))
        per_dir = []
        if peer_mode:
            fft.stage_timing(True)
            for direction in (("forward",) if fused_conv else ("forward", "backward")):
                barrier(ctx)
                if fused_conv:
                    fft.convolve_buffered(data_in, data_in, work, None, hf.scale.full)
                elif direction == "forward":
                    fft.forward_buffered(data_in, data_out, work, hf.scale.full)
                else:
                    fft.backward_buffered(data_out, data_in, work, hf.scale.none)
                torch.cuda.synchronize()
                per_dir.append([dict(direction=direction, stage=nm, ms=ms, local_bytes=lb, sent_bytes=sb) for nm, ms, lb, sb in fft.stage_times()])
            fft.stage_timing(False)
        # max over ranks of every stage time; the stage list has the same length on every rank (empty boxes emit zero-byte marks),
        # padded here all the same so that a mismatch can never hang the all_reduce
        flat = [e for d in per_dir for e in d]
        length = int(all_max(ctx, [float(len(flat))])[0])
        while len(flat) < length:
            flat.append(dict(direction="?", stage="(absent on rank 0)", ms=0.0, local_bytes=0, sent_bytes=0))
        if flat:
            for e, v in zip(flat, all_max(ctx, [e["ms"] for e in flat])):
                e["ms"] = v
            # bytes of the busiest rank of every stage (the plan may be uneven: tools/plan_traffic.py)
            both = all_max(ctx, [v for e in flat for v in (e["local_bytes"], e["sent_bytes"])])
            for i, e in enumerate(flat):
                e["local_bytes"], e["sent_bytes"] = int(both[2 * i]), int(both[2 * i + 1])
        multi = {"peer_memory": peer_mode, "stages": flat}

    # ---- roofline of the dominant kernel: the batched 1-D FFT pass, timed alone with CUDA events -----------------------
    roofline, stages = None, []
    peaks, peak_kind = measured_peaks()
    if rank == 0 and world_size == 1:
        from heffte_b200._lib import b200_fft1d_desc, b200_line_geom
        n0, n1, n2 = [int(v) for v in inbox.size]
        rsize = 4 if prec == 0 else 8
        csize = 2 * rsize
        # (kind, length, count_a, count_b, geometry in, geometry out, in buffer, out buffer, algorithmic bytes: one read + one write)
        if r2c:
            h = n0 // 2 + 1
            plans = [(1, n0, n1 * n2, 1, (1, n0, 0), (1, h, 0), data_in, data_out, n0 * n1 * n2 * rsize + h * n1 * n2 * csize),
                     (0, n1, h, n2, (h, 1, h * n1), (h, 1, h * n1), data_out, data_out, 2 * h * n1 * n2 * csize),
                     (0, n2, h * n1, 1, (h * n1, 1, 0), (h * n1, 1, 0), data_out, data_out, 2 * h * n1 * n2 * csize)]
        else:
            k, esize = (2, rsize) if r2r else (0, csize)
            plans = [(k, n0, n1 * n2, 1, (1, n0, 0), (1, n0, 0), data_in, data_in, 2 * n0 * n1 * n2 * esize),
                     (k, n1, n0, n2, (n0, 1, n0 * n1), (n0, 1, n0 * n1), data_in, data_in, 2 * n0 * n1 * n2 * esize),
                     (k, n2, n0 * n1, 1, (n0 * n1, 1, 0), (n0 * n1, 1, 0), data_in, data_in, 2 * n0 * n1 * n2 * esize)]
        for dim, (k, length, ca, cb, gi, go, src, dst, algo_bytes) in enumerate(plans):
            for direction in ((0, 1) if (r2c or r2r) else (0,)):
                d = b200_fft1d_desc(prec, k, length, ca, cb, b200_line_geom(*gi), b200_line_geom(*go))
                plan = ctypes.c_void_p()
                if lib.b200_fft1d_create(ctypes.byref(d), ctypes.byref(plan)) != 0:
                    continue
                pin, pout = ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr())
                if direction == 1:
                    pin, pout = pout, pin
                reps = 10
                events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
                for _ in range(3):
                    lib.b200_fft1d_execute(plan, direction, pin, pout, ctypes.c_double(1.0), None)
                torch.cuda.synchronize()
                if flush:
                    for a, b in events:
                        flush_l2(ctx)
                        a.record()
                        lib.b200_fft1d_execute(plan, direction, pin, pout, ctypes.c_double(1.0), None)
                        b.record()
                    torch.cuda.synchronize()
                    ms = sum(a.elapsed_time(b) for a, b in events) / reps
                else:
                    a, b = events[0]
                    a.record()
                    for _ in range(reps):
                        lib.b200_fft1d_execute(plan, direction, pin, pout, ctypes.c_double(1.0), None)
                    b.record()
                    torch.cuda.synchronize()
                    ms = a.elapsed_time(b) / reps
                name = lib.b200_fft1d_kernel_name(plan).decode()
                # the library takes the TMA-loaded variant of the 512-point fp64 strided kernel for rows at least 1 MiB apart
                # (csrc/fft_dispatch.cuh tma_tile_wanted; the choice is made per launch, the plan only knows the family)
                if name == "strided" and k == 0 and length == 512 and prec == 1 and gi[0] * csize >= (1 << 20) and ca % 8 == 0 \
                        and os.environ.get("HEFFTE_B200_TMA", "1")[:1] != "0":
                    name = "strided_tma"
                lib.b200_fft1d_destroy(plan)
                stages.append({"dim": dim, "direction": "backward" if direction else "forward", "kernel": name, "n": length, "ms": ms,
                               "GB/s": algo_bytes / ms * 1e-6, "algorithmic_bytes": algo_bytes,
                               "frac_of_%s_hbm" % peak_kind: algo_bytes / ms * 1e-6 / peaks["hbm_gbs"]})
        if stages:
            dominant = max(stages, key=lambda s: s["ms"])
            total_ms = sum(s["ms"] for s in stages)
            passes_per_step = 6.0
            algo_total = sum(s["algorithmic_bytes"] for s in stages) * (passes_per_step / len(stages))
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    traffic = json.load(f).get("fft_%s_kernel/%d/%s" % (dominant["kernel"], dominant["n"], "f64" if prec == 1 else "f32"))
            except Exception:
                pass
            roofline = {"bound": "hbm", "achieved": dominant["GB/s"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": dominant["GB/s"] / peaks["hbm_gbs"], "traffic": traffic,
                        "kernel": "fft_%s_kernel (dim %d, %s)" % (dominant["kernel"], dominant["dim"], dominant["direction"]),
                        "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6650 GB/s",
                        "algorithmic_bytes_per_launch": dominant["algorithmic_bytes"],
                        "whole_transform": {"algorithmic_GB_per_step": algo_total * 1e-9, "sum_of_passes_ms": total_ms,
                                            "measured_ms_per_step": ms_per_step,
                                            "frac_of_hbm_roofline": (algo_total / (peaks["hbm_gbs"] * 1e9)) / (ms_per_step * 1e-3)}}

    if rank == 0 and multi is not None and multi["stages"]:
        for e in multi["stages"]:
            if e["ms"] > 0:
                e["hbm_GB/s"] = e["local_bytes"] / e["ms"] * 1e-6
                e["nvlink_GB/s"] = e["sent_bytes"] / e["ms"] * 1e-6
        fwd = [e for e in multi["stages"] if e["direction"] == "forward"]
        work_stages = [e for e in fwd if e["stage"] != "fence" and e["stage"] != "start"]
        dominant = max(work_stages, key=lambda e: e["ms"]) if work_stages else None
        elem = (4 if prec == 0 else 8) * (1 if r2r else 2)
        d_bytes = float(max(nin, nout)) * elem                               # D: bytes of one rank's box
        hbm_bytes = 6.0 * d_bytes                                            # three passes, read + write
        nvl_bytes = float(sum(e["sent_bytes"] for e in fwd))
        t_hbm = hbm_bytes / (peaks["hbm_gbs"] * 1e9)
        t_nvl = nvl_bytes / (NVLINK_PEAK * 1e9)
        idle_ms = sum(e["ms"] for e in fwd if e["sent_bytes"] == 0 and e["stage"] != "start")
        if dominant is not None:
            sent_bound = dominant["sent_bytes"] / (NVLINK_PEAK * 1e9) >= dominant["local_bytes"] / (peaks["hbm_gbs"] * 1e9)
            roofline = {"bound": "nvlink" if sent_bound else "hbm",
                        "achieved": dominant["nvlink_GB/s"] if sent_bound else dominant["hbm_GB/s"],
                        "peak": NVLINK_PEAK if sent_bound else peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": (dominant["nvlink_GB/s"] / NVLINK_PEAK) if sent_bound else (dominant["hbm_GB/s"] / peaks["hbm_gbs"]),
                        "traffic": None, "kernel": dominant["stage"],
                        "peak_source": "measured peer copy 770 GB/s per direction (B200_PROFILING.md)" if sent_bound else peak_kind + " hbm_gbs",
                        "algorithmic_bytes_per_launch": dominant["sent_bytes"] if sent_bound else dominant["local_bytes"],
                        "whole_transform": {"hbm_algorithmic_GB": hbm_bytes * 1e-9, "nvlink_GB_sent_per_gpu": nvl_bytes * 1e-9,
                                            "t_hbm_ms": t_hbm * 1e3, "t_nvlink_ms": t_nvl * 1e3, "measured_ms": sec_per_transform * 1e3,
                                            "forward_ms_with_nvlink_idle": idle_ms,
                                            "frac_of_overlap_roofline": max(t_hbm, t_nvl) / sec_per_transform,
                                            "frac_of_serial_roofline": (t_hbm + t_nvl) / sec_per_transform}}
        stages = multi["stages"]

    # ---- CPU baseline: the unmodified reference on this box's host cores, bounded sample ------------------------------
    cpu = None
    if rank == 0 and world_size == 1 and cpu_wanted and not args.no_cpu_baseline and not conv:
        csize_ = args.cpu_size or size
        r = run_reference_speed3d(kind, precision, csize_, 1)
        if r is not None and "error" not in r:
            cpu = {"value": r["gflops"], "unit": "GFlop/s", "cores": r["ranks"], "kind": "reference", "sample": r["sample"],
                   "host_cores_available": r["cores"]}
        else:
            cpu = {"value": None, "unit": "GFlop/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % ((r or {}).get("error", "oracle/_ref not built"))}

    result = None
    if rank == 0:
        grid = "x".join(str(v) for v in in_grid)
        result = {
            "metric": "speed3d_%s GFlop/s (5*N*log2(N)/t)" % kind, "value": value, "unit": "GFlop/s", "n_gpus": world_size,
            "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if prec == 1 else "f32", "data": "synthetic",
            "config": {"workload": workload_name(kind, precision, n),
                       "layout": "bricks %s%s, %s, %s, in-place, step = forward(scale full)+backward" % (
                           grid, " (pencil-shaped in/out)" if io_pencils else "", "reorder" if reorder else "no-reorder", "slabs" if slabs else ("pencils" if pencils else "decomposition chosen by the planner")),
                       "batch": batch,
                       "registered_arrays": registered,
                       "l2": ("working set %.0f MB per GPU exceeds the 126 MB L2" % working_set_mb) if not flush else
                             ("working set %.0f MB per GPU: a 512 MB buffer is overwritten between the timed steps (flush time not counted)" % working_set_mb),
                       "l2_slab_mb": os.environ.get("HEFFTE_B200_L2_SLAB_MB"),
                       "executed_plan": executed,
                       "comm": ("peer memory: NVLink stores fused into the FFT kernels" if (multi and multi["peer_memory"]) else
                                ("peer memory" if fft.uses_peer_memory(prec) else "nccl send/recv")) if distributed else "none"},
            "max_roundtrip_error": None if conv else err,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": e2e,
            "roofline": roofline,
            "stages": stages,
            "cpu_baseline": cpu,
        }
        if parity_info is not None:
            result.update({"parity_rel_l2": parity_info.get("parity_rel_l2"), "parity_ok": parity_info.get("parity_ok"), "parity": parity_info})
    del fft, data_in, data_out, reference_copy, work
    torch.cuda.empty_cache()
    return result


def secondary_list(args, world_size):
    """the other BASELINE.json configurations, each a short run of the same harness"""
    items = [dict(kind="r2c", size=(512, 512, 512), precision="double"),
             dict(kind="r2r", size=(512, 512, 512), precision="double"),
             dict(kind="conv", size=(512, 512, 512), precision="double")]
    if world_size == 1:
        items.append(dict(kind="c2c", size=(256, 256, 256), precision="float", force_flush=True))
    if world_size >= 8:
        # BASELINE cfg4: speed3d_c2c single 1024^3 -reorder, slabs against pencils
        items.append(dict(kind="c2c", size=(1024, 1024, 1024), precision="float", reorder=True, slabs=True))
        items.append(dict(kind="c2c", size=(1024, 1024, 1024), precision="float", reorder=True, pencils=True))
        items.append(dict(kind="c2c", size=(512, 512, 512), precision="double", io_pencils=True))
    return items


def b200_arm(args):
    if args.l2_slab_mb is not None:
        os.environ["HEFFTE_B200_L2_SLAB_MB"] = str(args.l2_slab_mb)
    ctx = make_context(args)
    line = run_workload(ctx, args, args.kind, args.size, args.precision, args.steps, args.warmup, primary=True,
                        reorder=args.reorder, slabs=args.slabs, pencils=args.pencils, io_pencils=args.io_pencils, batch=args.batch)
    default_workload = (args.kind == "c2c" and tuple(args.size) == (512, 512, 512) and args.precision == "double" and args.batch == 1
                        and not (args.reorder or args.slabs or args.pencils or args.io_pencils))
    if default_workload and not args.no_secondary:
        secondary = []
        for item in secondary_list(args, ctx.world_size):
            try:
                r = run_workload(ctx, args, item["kind"], item["size"], item["precision"], max(3, min(args.steps, 5)), 3, primary=False,
                                 reorder=item.get("reorder", False), io_pencils=item.get("io_pencils", False),
                                 slabs=item.get("slabs", False), pencils=item.get("pencils", False),
                                 force_flush=item.get("force_flush", False), e2e_wanted=False, cpu_wanted=False)
            except Exception as e:  # noqa: BLE001  (a secondary configuration must not take the headline down)
                r = {"config": {"workload": workload_name(item["kind"], item["precision"], item["size"])}, "error": repr(e)} if ctx.rank == 0 else None
                if ctx.distributed:
                    raise
            if ctx.rank == 0 and r is not None:
                keep = {k: r.get(k) for k in ("metric", "value", "unit", "ms_per_step", "dtype", "max_roundtrip_error", "parity_rel_l2", "parity_ok", "error") if k in r}
                keep["workload"] = r["config"]["workload"]
                keep["layout"] = r["config"].get("layout")
                keep["l2"] = r["config"].get("l2")
                if r.get("roofline"):
                    rf = r["roofline"]
                    keep["roofline"] = {k: rf.get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "kernel")}
                    keep["roofline"]["whole_transform"] = rf.get("whole_transform")
                keep["stages"] = [{k: s.get(k) for k in ("dim", "direction", "kernel", "n", "ms", "GB/s", "stage", "nvlink_GB/s", "hbm_GB/s") if k in s}
                                  for s in r.get("stages", []) if s.get("stage") not in ("start",)]
                secondary.append(keep)
        if ctx.rank == 0:
            line["secondary"] = secondary
    if ctx.rank == 0:
        print(json.dumps(line))
    if ctx.distributed:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # started without a launcher: one process per GPU under torch.distributed.run, same arguments
        import socket
        with socket.socket() as probe:
            probe.bind(("127.0.0.1", 0))
            port = probe.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
