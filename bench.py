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
measured HBM peak, `cpu_baseline` = the unmodified reference (stock backend, threads-as-ranks MPI stand-in) timed on
this box's host cores.
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
    ap.add_argument("--slabs", action="store_true")
    ap.add_argument("--io-pencils", action="store_true", help="pencil-shaped in/out boxes (speed3d -io_pencils)")
    ap.add_argument("--l2-slab-mb", type=float, default=None,
                    help="experimental: run pairs of local transforms slab by slab through the L2 cache (sets HEFFTE_B200_L2_SLAB_MB)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-size", type=int, nargs=3, default=None, help="size of the CPU sample (default: the workload itself)")
    return ap.parse_args()


def gflops(n, seconds_per_transform):
    N = float(n[0]) * n[1] * n[2]
    return 5.0 * N * math.log2(N) * 1e-9 / seconds_per_transform


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
    binary = ref_lib.binary("speed3d_" + kind)
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
            "cores": cores, "wall_s": wall, "isa": os.path.basename(os.path.dirname(binary)),
            "sample": "speed3d_%s stock %s %dx%dx%d -n%d on %d thread-ranks (reference compiled in place, %s)" % (
                kind, precision, size[0], size[1], size[2], nruns, ranks, os.path.basename(os.path.dirname(binary)))}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size = args.cpu_size or args.size
    # each "step" is one bounded sample: the reference benchmark's own timed loop of one forward+backward pair
    nruns = 1
    results = []
    for _ in range(max(1, min(args.steps, 2))):
        r = run_reference_speed3d(args.kind, args.precision, size, nruns)
        if r is None or "error" in r:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref speed3d binary missing or failed: %s" % (r or {}).get("error", "not built")}))
            return
        results.append(r)
    best = max(results, key=lambda r: r["gflops"])
    line = {
        "impl": "reference", "metric": "speed3d_%s GFlop/s (5*N*log2(N)/t)" % args.kind, "value": best["gflops"], "unit": "GFlop/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 2e3 * best["seconds_per_transform"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if args.precision == "double" else "f32",
        "data": "synthetic",
        "config": {"workload": "speed3d_%s %s %dx%dx%d" % (args.kind, args.precision, size[0], size[1], size[2]),
                   "backend": "stock (FFTW and MPI are absent from the image)", "ranks": best["ranks"]},
        "cpu_baseline": {"value": best["gflops"], "unit": "GFlop/s", "cores": best["ranks"], "kind": "reference", "sample": best["sample"]},
        "e2e": {"value": best["gflops"], "unit": "GFlop/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# the b200 arm
# ----------------------------------------------------------------------------------------------------------------
def b200_arm(args):
    if args.l2_slab_mb is not None:
        os.environ["HEFFTE_B200_L2_SLAB_MB"] = str(args.l2_slab_mb)
    import numpy as np
    import torch
    import heffte_b200 as hf
    from heffte_b200 import _lib, build
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the library is built in-tree ahead of time (python -m heffte_b200.build / __graft_entry__.build()); build here only if it is
    # missing, and never from several ranks at once
    if not os.path.exists(build.library_path()):
        if world_size > 1:
            raise SystemExit("bench.py: %s is missing; run `python -m heffte_b200.build` before a multi-rank launch" % build.library_path())
        build.build_library()
    lib = _lib.load()

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    distributed = world_size > 1
    if distributed:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = hf.comm_from_torch()
    else:
        dist = None
        comm = hf.comm_self()

    n = tuple(args.size)
    prec = 0 if args.precision == "float" else 1
    rdtype = torch.float32 if prec == 0 else torch.float64
    cdtype = torch.complex64 if prec == 0 else torch.complex128
    world = hf.box3d((0, 0, 0), (n[0] - 1, n[1] - 1, n[2] - 1))

    if args.io_pencils:
        g2 = hf.heffte.make_procgrid(world_size)
        in_grid, out_grid = [1, g2[0], g2[1]], [g2[0], g2[1], 1]
    else:
        in_grid = out_grid = hf.heffte.proc_setup_min_surface(world, world_size)
    r2c = args.kind == "r2c"
    inbox = hf.heffte.split_world(world, in_grid)[rank]
    if r2c:
        cworld = hf.box3d((0, 0, 0), (n[0] // 2, n[1] - 1, n[2] - 1))
        outbox = hf.heffte.split_world(cworld, out_grid)[rank]
    else:
        outbox = hf.heffte.split_world(world, out_grid)[rank]

    r2r = args.kind == "r2r"
    tag = hf.backend.b200_cos if r2r else hf.backend.b200
    options = hf.plan_options(tag, use_reorder=args.reorder, use_pencils=not args.slabs)
    fft = hf.fft3d_r2c(tag, inbox, outbox, 0, comm, options) if r2c else hf.fft3d(tag, inbox, outbox, comm, options)

    # the plan that runs (pure host planning, csrc/plan_logic.h): process grids of input, the three transform stages, output
    executed = None
    try:
        world_boxes_in = hf.heffte.split_world(world, in_grid)
        world_boxes_out = hf.heffte.split_world(cworld if r2c else world, out_grid)
        shapes, _, swaps = hf.heffte.execution_plan(world_boxes_in, world_boxes_out, r2c_direction=0 if r2c else -1,
                                                    use_reorder=bool(args.reorder or r2r), use_pencils=not args.slabs)

        def grid_of(boxes):
            return "x".join(str(len({(b[d], b[3 + d]) for b in boxes if all(b[3 + k] >= b[k] for k in range(3))})) for d in range(3))
        executed = {"grids": " -> ".join(grid_of(shapes[i]) for i in (0, 4, 5, 6, 7)), "refinements_applied": swaps}
    except Exception as e:  # noqa: BLE001  (reporting only)
        executed = {"error": repr(e)}

    gen = torch.Generator(device="cuda")
    gen.manual_seed(4242 + rank)
    nin, nout = fft.size_inbox(), fft.size_outbox()
    if r2c:
        data_in = torch.rand(nin, dtype=rdtype, device="cuda", generator=gen)
        data_out = torch.empty(nout, dtype=cdtype, device="cuda")
    elif r2r:
        data_in = torch.rand(max(nin, nout), dtype=rdtype, device="cuda", generator=gen)
        data_out = data_in
    else:
        # speed3d: complex data with zero imaginary part, transformed in place
        data_in = torch.complex(torch.rand(max(nin, nout), dtype=rdtype, device="cuda", generator=gen),
                                torch.zeros(max(nin, nout), dtype=rdtype, device="cuda"))
        data_out = data_in
    reference_copy = data_in.clone()
    work = torch.empty(fft.size_workspace(), dtype=rdtype if r2r else cdtype, device="cuda")

    conv = args.kind == "conv"

    def step():
        fft.forward_buffered(data_in, data_out, work, hf.scale.full)
        if conv:
            data_out.mul_(data_out)      # the caller's pointwise product in spectral space (benchmarks/convolution.cpp:89-94)
        fft.backward_buffered(data_out, data_in, work, hf.scale.none)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # accuracy of the round trip (benchmarks/speed3d.h:213-220)
    err = float((data_in - reference_copy).abs().max().item()) if not conv else float("nan")
    data_in.copy_(reference_copy)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.b200_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    launches = lib.b200_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if distributed:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        e = torch.tensor([err], dtype=torch.float64, device="cuda")
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        err = float(e.item())
    ms_per_step = elapsed_ms / args.steps
    sec_per_transform = ms_per_step * 1e-3 / 2.0
    value = gflops(n, sec_per_transform)

    # ---- end to end: host buffers in, host buffers out, through the public plan API ---------------------------------
    e2e = None
    if not args.no_e2e and not conv:
        real_bytes = 4 if prec == 0 else 8
        host_in = torch.empty(nin if r2c else max(nin, nout), dtype=rdtype if (r2c or r2r) else cdtype).pin_memory()
        host_mid = torch.empty(nout if r2c else max(nin, nout), dtype=rdtype if r2r else cdtype).pin_memory()
        host_in.copy_(reference_copy.cpu())
        np_in, np_mid = host_in.numpy(), host_mid.numpy()
        np_back = torch.empty_like(host_in).pin_memory().numpy()
        e2e_steps = max(2, min(args.steps, 5))

        def e2e_step():
            fft.forward(np_in, np_mid, hf.scale.full)      # H2D(in) + forward + D2H(out)
            fft.backward(np_mid, np_back, hf.scale.none)   # H2D(out) + backward + D2H(in)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        if distributed:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        in_bytes = host_in.numel() * host_in.element_size()
        mid_bytes = host_mid.numel() * host_mid.element_size()
        e2e = {"value": gflops(n, e2e_s / 2.0), "unit": "GFlop/s", "h2d_bytes_per_step": in_bytes + mid_bytes,
               "d2h_bytes_per_step": in_bytes + mid_bytes, "steps": e2e_steps,
               "note": "per rank bytes; step = forward(host in -> host out) + backward(host out -> host in), pinned buffers"}

    # ---- multi-GPU: device time of every stage of one forward and one backward transform (CUDA events on the plan's stream) ----
    multi = None
    if distributed:
        peer_mode = bool(fft.uses_peer_memory(prec))
        per_dir = []
        if peer_mode:
            fft.stage_timing(True)
            for direction in ("forward", "backward"):
                barrier()
                if direction == "forward":
                    fft.forward_buffered(data_in, data_out, work, hf.scale.full)
                else:
                    fft.backward_buffered(data_out, data_in, work, hf.scale.none)
                torch.cuda.synchronize()
                per_dir.append([dict(direction=direction, stage=nm, ms=ms, local_bytes=lb, sent_bytes=sb) for nm, ms, lb, sb in fft.stage_times()])
            fft.stage_timing(False)
        # max over ranks of every stage time (same stage list on every rank)
        flat = [e for d in per_dir for e in d]
        if flat:
            t = torch.tensor([e["ms"] for e in flat], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            for e, v in zip(flat, t.tolist()):
                e["ms"] = v
            # bytes of the busiest rank of every stage (the plan may be uneven: tools/plan_traffic.py)
            t = torch.tensor([[e["local_bytes"], e["sent_bytes"]] for e in flat], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            for e, (lb, sb) in zip(flat, t.tolist()):
                e["local_bytes"], e["sent_bytes"] = int(lb), int(sb)
        multi = {"peer_memory": peer_mode, "stages": flat}

    # ---- roofline of the dominant kernel: the batched 1-D FFT pass, timed alone with CUDA events -----------------------
    roofline, stages = None, []
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        from heffte_b200._lib import b200_fft1d_desc, b200_line_geom
        # the local box of the first FFT stage on this rank; on one GPU this is the whole world
        stage_box = [int(v) for v in inbox.size]
        if world_size > 1:
            stage_box = None  # stage boxes differ per stage; report per-stage numbers only on one GPU
        if stage_box is not None and not conv:
            n0, n1, n2 = stage_box
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
                d = b200_fft1d_desc(prec, k, length, ca, cb, b200_line_geom(*gi), b200_line_geom(*go))
                plan = ctypes.c_void_p()
                if lib.b200_fft1d_create(ctypes.byref(d), ctypes.byref(plan)) != 0:
                    continue
                pin, pout = ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr())
                for _ in range(3):
                    lib.b200_fft1d_execute(plan, 0, pin, pout, ctypes.c_double(1.0), None)
                torch.cuda.synchronize()
                reps = 10
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    lib.b200_fft1d_execute(plan, 0, pin, pout, ctypes.c_double(1.0), None)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / reps
                name = lib.b200_fft1d_kernel_name(plan).decode()
                lib.b200_fft1d_destroy(plan)
                stages.append({"dim": dim, "kernel": name, "n": length, "ms": ms, "GB/s": algo_bytes / ms * 1e-6, "algorithmic_bytes": algo_bytes,
                               "frac_of_%s_hbm" % peak_kind: algo_bytes / ms * 1e-6 / peaks["hbm_gbs"]})
            if stages:
                dominant = max(stages, key=lambda s: s["ms"])
                total_ms = sum(s["ms"] for s in stages)
                traffic = None
                try:
                    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                        traffic = json.load(f).get("fft_%s_kernel/%d/%s" % (dominant["kernel"], dominant["n"], "f64" if prec == 1 else "f32"))
                except Exception:
                    pass
                roofline = {"bound": "hbm", "achieved": dominant["GB/s"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": dominant["GB/s"] / peaks["hbm_gbs"], "traffic": traffic,
                            "kernel": "fft_%s_kernel (dim %d)" % (dominant["kernel"], dominant["dim"]),
                            "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6650 GB/s",
                            "algorithmic_bytes_per_launch": dominant["algorithmic_bytes"],
                            "whole_transform": {"algorithmic_GB": sum(s["algorithmic_bytes"] for s in stages) * 1e-9, "sum_of_passes_ms": total_ms,
                                                "measured_ms": sec_per_transform * 1e3,
                                                "frac_of_hbm_roofline": (sum(s["algorithmic_bytes"] for s in stages) / (peaks["hbm_gbs"] * 1e9)) / sec_per_transform}}

    if rank == 0 and multi is not None and multi["stages"]:
        peaks, peak_kind = measured_peaks()
        nvlink_peak = 770.0   # GB/s per direction per GPU: measured peer copy on this pool (B200_PROFILING.md; tools/ipc_probe.py saw 760)
        for e in multi["stages"]:
            if e["ms"] > 0:
                e["hbm_GB/s"] = e["local_bytes"] / e["ms"] * 1e-6
                e["nvlink_GB/s"] = e["sent_bytes"] / e["ms"] * 1e-6
        fwd = [e for e in multi["stages"] if e["direction"] == "forward"]
        work_stages = [e for e in fwd if e["stage"] != "fence" and e["stage"] != "start"]
        dominant = max(work_stages, key=lambda e: e["ms"]) if work_stages else None
        elem = (8 if prec == 0 else 16)
        d_bytes = float(max(nin, nout)) * elem                               # D: bytes of one rank's box
        hbm_bytes = 6.0 * d_bytes                                            # three passes, read + write
        nvl_bytes = float(sum(e["sent_bytes"] for e in fwd))
        t_hbm = hbm_bytes / (peaks["hbm_gbs"] * 1e9)
        t_nvl = nvl_bytes / (nvlink_peak * 1e9)
        if dominant is not None:
            sent_bound = dominant["sent_bytes"] / (nvlink_peak * 1e9) >= dominant["local_bytes"] / (peaks["hbm_gbs"] * 1e9)
            roofline = {"bound": "nvlink" if sent_bound else "hbm",
                        "achieved": dominant["nvlink_GB/s"] if sent_bound else dominant["hbm_GB/s"],
                        "peak": nvlink_peak if sent_bound else peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": (dominant["nvlink_GB/s"] / nvlink_peak) if sent_bound else (dominant["hbm_GB/s"] / peaks["hbm_gbs"]),
                        "traffic": None, "kernel": dominant["stage"],
                        "peak_source": "measured peer copy 770 GB/s per direction (B200_PROFILING.md)" if sent_bound else peak_kind + " hbm_gbs",
                        "algorithmic_bytes_per_launch": dominant["sent_bytes"] if sent_bound else dominant["local_bytes"],
                        "whole_transform": {"hbm_algorithmic_GB": hbm_bytes * 1e-9, "nvlink_GB_sent_per_gpu": nvl_bytes * 1e-9,
                                            "t_hbm_ms": t_hbm * 1e3, "t_nvlink_ms": t_nvl * 1e3, "measured_ms": sec_per_transform * 1e3,
                                            "frac_of_overlap_roofline": max(t_hbm, t_nvl) / sec_per_transform,
                                            "frac_of_serial_roofline": (t_hbm + t_nvl) / sec_per_transform}}
        stages = multi["stages"]

    # ---- CPU baseline: the unmodified reference on this box's host cores, bounded sample ------------------------------
    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline and not conv:
        size = args.cpu_size or args.size
        r = run_reference_speed3d(args.kind, args.precision, size, 1)
        if r is not None and "error" not in r:
            cpu = {"value": r["gflops"], "unit": "GFlop/s", "cores": r["ranks"], "kind": "reference", "sample": r["sample"],
                   "host_cores_available": r["cores"]}
        else:
            cpu = {"value": None, "unit": "GFlop/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % ((r or {}).get("error", "oracle/_ref not built"))}

    if rank == 0:
        grid = "x".join(str(v) for v in in_grid)
        working_set_mb = max(nin, nout) * ((4 if prec == 0 else 8) * (1 if r2r else 2)) / 1e6
        line = {
            "metric": "speed3d_%s GFlop/s (5*N*log2(N)/t)" % args.kind, "value": value, "unit": "GFlop/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if prec == 1 else "f32", "data": "synthetic",
            "config": {"workload": "speed3d_%s %s %dx%dx%d, bricks %s, %s, %s, in-place, step = forward(scale full)+backward" % (
                           args.kind, args.precision, n[0], n[1], n[2], grid, "reorder" if args.reorder else "no-reorder",
                           "slabs" if args.slabs else "pencils"),
                       "l2": ("working set %.0f MB per GPU exceeds the 126 MB L2" if working_set_mb > 126 else
                              "working set %.0f MB per GPU FITS in the 126 MB L2 and nothing flushes it: not a valid bench configuration") % working_set_mb,
                       "l2_slab_mb": os.environ.get("HEFFTE_B200_L2_SLAB_MB"),
                       "executed_plan": executed,
                       "comm": ("peer memory: NVLink stores fused into the FFT kernels" if (multi and multi["peer_memory"]) else "nccl send/recv") if distributed else "none"},
            "max_roundtrip_error": None if conv else err,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": e2e,
            "roofline": roofline,
            "stages": stages,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


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
