"""
TEST INFRASTRUCTURE ONLY.  ctypes loader for oracle/_ref/<isa>/libheffte_ref.so -- the unmodified
reference (stock backend) compiled in place by oracle/Makefile through oracle/ref_shim.cpp.
Used by tests/, tests/golden/make_golden.py and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")

KIND = {"c2c": 0, "r2c": 1, "cos": 2, "sin": 3, "cos1": 4}
SCALE = {"none": 0, "full": 1, "symmetric": 2}


def cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def isa_dir():
    """Pick the widest prebuilt flavour this host can execute."""
    flags = cpu_flags()
    order = []
    if "avx512f" in flags and "avx512dq" in flags:
        order.append("avx512")
    if "avx2" in flags and "fma" in flags:
        order.append("avx2")
    for name in order:
        d = os.path.join(REF_DIR, name)
        if os.path.exists(os.path.join(d, ".done")):
            return d
    return None


def build(verbose=False):
    """Build oracle/_ref when the reference tree is present (this container); no-op on the GPU box."""
    if not os.path.exists("/root/reference/include/heffte.h"):
        return False
    out = subprocess.run(["make", "-C", _HERE, "ref"], capture_output=not verbose, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle/_ref build failed:\n" + (out.stderr or ""))
    return True


def available():
    return isa_dir() is not None


_lib = None


def lib():
    global _lib
    if _lib is None:
        d = isa_dir()
        if d is None:
            raise RuntimeError("oracle/_ref is not built (run `make -C oracle` where /root/reference exists)")
        _lib = ctypes.CDLL(os.path.join(d, "libheffte_ref.so"))
        _lib.ref_make_data.argtypes = [ctypes.c_longlong, ctypes.c_void_p]
        _lib.ref_make_data.restype = None
    return _lib


def binary(name):
    d = isa_dir()
    return None if d is None else os.path.join(d, name)


def _i32(values):
    return np.ascontiguousarray(np.asarray(values, dtype=np.int32).reshape(-1))


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def make_data(count):
    out = np.empty(count, dtype=np.float64)
    lib().ref_make_data(count, _ptr(out))
    return out


def _ctype(prec):
    return np.complex64 if prec == 0 else np.complex128


def _rtype(prec):
    return np.float32 if prec == 0 else np.float64


def exec1d_c2c(data, box, dim, backward=False, prec=1):
    buf = np.ascontiguousarray(np.asarray(data).astype(_ctype(prec)).reshape(-1))
    b = _i32(box.nine())
    rc = lib().ref_exec1d_c2c(prec, _ptr(b), dim, int(backward), _ptr(buf))
    assert rc == 0
    return buf


def exec1d_r2c(data, box, dim, prec=1):
    rbuf = np.ascontiguousarray(np.asarray(data).astype(_rtype(prec)).reshape(-1))
    cbuf = np.zeros(box.r2c(dim).count(), dtype=_ctype(prec))
    b = _i32(box.nine())
    rc = lib().ref_exec1d_r2c(prec, _ptr(b), dim, 0, _ptr(rbuf), _ptr(cbuf))
    assert rc == 0
    return cbuf


def exec1d_c2r(data, box, dim, prec=1):
    cbuf = np.ascontiguousarray(np.asarray(data).astype(_ctype(prec)).reshape(-1))
    rbuf = np.zeros(box.count(), dtype=_rtype(prec))
    b = _i32(box.nine())
    rc = lib().ref_exec1d_r2c(prec, _ptr(b), dim, 1, _ptr(rbuf), _ptr(cbuf))
    assert rc == 0
    return rbuf


def exec1d_r2r(data, box, dim, kind, backward=False, prec=1):
    buf = np.ascontiguousarray(np.asarray(data).astype(_rtype(prec)).reshape(-1))
    b = _i32(box.nine())
    rc = lib().ref_exec1d_r2r(prec, KIND[kind], _ptr(b), dim, int(backward), _ptr(buf))
    assert rc == 0
    return buf


def fft3d(kind, prec, inboxes, outboxes, inputs, backward=False, scaling="none", r2c_dir=0,
          use_reorder=True, algorithm=0, use_pencils=True, subranks=-1):
    """
    Run heffte::fft3d / fft3d_r2c <stock> on len(inboxes) thread-ranks.  `inputs` is a list of per-rank
    flat arrays (or None to only query the workspace sizes).  Returns (outputs, workspace_sizes).
    """
    n = len(inboxes)
    ib = _i32([b.nine() for b in inboxes])
    ob = _i32([b.nine() for b in outboxes])
    ws = np.zeros(n, dtype=np.int64)
    real_in = (kind == "r2c" and not backward) or kind in ("cos", "sin", "cos1")
    real_out = (kind == "r2c" and backward) or kind in ("cos", "sin", "cos1")
    ins, outs = [], []
    in_ptrs = out_ptrs = None
    if inputs is not None:
        for r in range(n):
            dt_in = _rtype(prec) if real_in else _ctype(prec)
            dt_out = _rtype(prec) if real_out else _ctype(prec)
            ins.append(np.ascontiguousarray(np.asarray(inputs[r]).astype(dt_in).reshape(-1)))
            count = inboxes[r].count() if backward else outboxes[r].count()
            outs.append(np.zeros(max(count, 1), dtype=dt_out))
        in_ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in ins])
        out_ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in outs])
    rc = lib().ref_fft3d(KIND[kind], prec, n, _ptr(ib), _ptr(ob), r2c_dir, int(backward), SCALE[scaling],
                         int(use_reorder), algorithm, int(use_pencils), subranks,
                         in_ptrs, out_ptrs, _ptr(ws))
    assert rc == 0, "reference fft3d failed"
    if inputs is not None:
        for r in range(n):
            count = inboxes[r].count() if backward else outboxes[r].count()
            outs[r] = outs[r][:count]
    return outs, ws


def plan_operations(inboxes, outboxes, r2c_dir=-1, use_reorder=True, algorithm=0, use_pencils=True, subranks=-1, rank=0):
    """Returns (shapes[8][nranks] as lists of 9 ints, fft_direction, index_count)."""
    n = len(inboxes)
    ib = _i32([b.nine() for b in inboxes])
    ob = _i32([b.nine() for b in outboxes])
    shapes = np.zeros(8 * n * 9, dtype=np.int32)
    fdir = np.zeros(3, dtype=np.int32)
    count = ctypes.c_longlong(0)
    rc = lib().ref_plan_operations(n, _ptr(ib), _ptr(ob), r2c_dir, int(use_reorder), algorithm, int(use_pencils),
                                   subranks, rank, _ptr(shapes), _ptr(fdir), ctypes.byref(count))
    assert rc == 0, "reference plan_operations failed"
    return shapes.reshape(8, n, 9).tolist(), fdir.tolist(), count.value


def make_procgrid(nprocs):
    g = np.zeros(2, dtype=np.int32)
    lib().ref_make_procgrid(nprocs, _ptr(g))
    return g.tolist()


def proc_setup_min_surface(world, nprocs):
    g = np.zeros(3, dtype=np.int32)
    w = _i32(world.nine())
    lib().ref_proc_setup_min_surface(_ptr(w), nprocs, _ptr(g))
    return g.tolist()


def split_world(world, grid):
    n = grid[0] * grid[1] * grid[2]
    out = np.zeros(9 * n, dtype=np.int32)
    w = _i32(world.nine())
    g = _i32(grid)
    lib().ref_split_world(_ptr(w), _ptr(g), _ptr(out))
    return out.reshape(n, 9).tolist()


def overlap_map_transpose(me, destination, boxes):
    n = len(boxes)
    proc = np.zeros(n, dtype=np.int32)
    offset = np.zeros(n, dtype=np.int32)
    sizes = np.zeros(n, dtype=np.int32)
    plans = np.zeros(10 * n, dtype=np.int32)
    d = _i32(destination.nine())
    b = _i32([x.nine() for x in boxes])
    k = lib().ref_overlap_map_transpose(me, n, _ptr(d), _ptr(b), _ptr(proc), _ptr(offset), _ptr(sizes), _ptr(plans))
    return [dict(proc=int(proc[i]), offset=int(offset[i]), size=int(sizes[i]), plan=plans[10 * i:10 * i + 10].tolist())
            for i in range(k)]


def pack(plan10, mode, src, dst):
    """mode 0 direct pack, 1 direct unpack, 2 transpose unpack; src/dst numpy arrays of 4/8/16-byte items."""
    p = _i32(plan10)
    rc = lib().ref_pack(src.dtype.itemsize, _ptr(p), mode, _ptr(src), _ptr(dst))
    assert rc == 0
