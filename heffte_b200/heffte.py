"""
Python front-end of the b200 backend: the same module-level API as the reference's ctypes binding
(reference python/heffte.py:109-407 -- backend, scale, box3d, fft3d(), fft3d_r2c(), plan.forward/backward,
size_inbox/outbox/workspace), bound to libheffte_b200.so.

Differences forced by the platform: there is no MPI / mpi4py in the image, so `comm` is a `heffte_b200.communicator`
(comm_self() or comm_from_torch() over torch.distributed + NCCL) instead of an mpi4py communicator; device arrays
are torch CUDA tensors (the reference takes numba device arrays); numpy arrays take the host path
(pinned staging + H2D/D2H inside the call).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import heffte_plan_options, LP_plan


class heffte_input_error(Exception):
    def __init__(self, message):
        self.message = message
        super().__init__(message)


class backend:
    # reference python/heffte.py:109-115 plus the new ids (include/heffte_b200.h)
    stock = 0
    fftw = 1
    mkl = 2
    cufft = 10
    rocm = 11
    b200 = 12
    b200_cos = 13
    b200_sin = 14
    b200_cos1 = 15
    valid = [12, 13, 14, 15]


class scale:
    none = 0
    full = 1
    symmetric = 2


class reshape_algorithm:
    alltoallv = 0
    p2p_plined = 1
    p2p = 2
    alltoall = 3


class box3d:
    """reference python/heffte.py:122-131 / include/heffte_geometry.h:67-133 (inclusive low/high, order)."""

    def __init__(self, clow, chigh, corder=(0, 1, 2)):
        self.low = np.array(clow, dtype=np.int32)
        self.high = np.array(chigh, dtype=np.int32)
        self.order = np.array(corder, dtype=np.int32)
        assert self.low.size == 3 and self.high.size == 3 and self.order.size == 3
        assert sorted(int(v) for v in self.order) == [0, 1, 2]

    @property
    def size(self):
        return self.high - self.low + 1

    def count(self):
        s = self.size
        return 0 if (s <= 0).any() else int(s[0]) * int(s[1]) * int(s[2])

    def empty(self):
        return self.count() == 0

    def nine(self):
        return [int(v) for v in self.low] + [int(v) for v in self.high] + [int(v) for v in self.order]

    def __repr__(self):
        return "box3d(%s, %s, %s)" % (self.low.tolist(), self.high.tolist(), self.order.tolist())


class plan_options:
    """reference include/heffte_plan_logic.h:131-176"""

    def __init__(self, backend_tag=backend.b200, use_reorder=None, algorithm=reshape_algorithm.alltoallv, use_pencils=None, use_gpu_aware=True):
        self.use_reorder = (backend_tag != backend.b200) if use_reorder is None else bool(use_reorder)
        self.algorithm = algorithm
        # True / False: that decomposition is executed, as in the reference; None: the planner picks the one that moves less over NVLink
        self.use_pencils = None if use_pencils is None else bool(use_pencils)
        self.use_gpu_aware = bool(use_gpu_aware)
        self.num_subranks = -1

    def use_subcomm(self, num_subranks):
        """reference include/heffte_plan_logic.h:100-129 (C++ only there): intermediate stages on the first num_subranks ranks"""
        self.num_subranks = int(num_subranks)

    def as_struct(self):
        return heffte_plan_options(int(self.use_reorder), int(self.algorithm), 2 if self.use_pencils is None else int(self.use_pencils), int(self.use_gpu_aware))


# ----------------------------------------------------------------------------------------------------------------
# communicators
# ----------------------------------------------------------------------------------------------------------------
class communicator:
    def __init__(self, handle, keepalive=None):
        self.handle = handle
        self._keepalive = keepalive

    def rank(self):
        return _lib.load().heffte_comm_rank(self.handle)

    def size(self):
        return _lib.load().heffte_comm_size(self.handle)

    def __del__(self):
        try:
            if self.handle:
                _lib.load().heffte_comm_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def comm_self():
    lib = _lib.load()
    handle = ctypes.c_void_p()
    if lib.heffte_comm_create_self(ctypes.byref(handle)) != 0:
        raise heffte_input_error(_lib.last_error())
    return communicator(handle)


def comm_threads(size, devices=None):
    """`size` ranks in this process, one host thread per rank (rank r on CUDA device devices[r], default all on device 0)."""
    lib = _lib.load()
    handles = (ctypes.c_void_p * size)()
    dev = None
    if devices is not None:
        dev = (ctypes.c_int * size)(*[int(d) for d in devices])
    if lib.heffte_comm_create_threads(size, dev, handles) != 0:
        raise heffte_input_error(_lib.last_error())
    return [communicator(ctypes.c_void_p(h)) for h in handles]


def comm_from_torch(group=None, device=None):
    """One rank per GPU: share an NCCL unique id through torch.distributed, then build the NCCL communicator."""
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    rank, size = dist.get_rank(group), dist.get_world_size(group)
    if device is not None:
        torch.cuda.set_device(device)
    ident = (ctypes.c_char * 128)()
    if rank == 0 and lib.heffte_comm_nccl_unique_id(ident) != 0:
        raise heffte_input_error(_lib.last_error())
    payload = [bytes(ident)]
    dist.broadcast_object_list(payload, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ident = (ctypes.c_char * 128).from_buffer_copy(payload[0])
    handle = ctypes.c_void_p()
    if lib.heffte_comm_create_nccl(rank, size, ident, ctypes.byref(handle)) != 0:
        raise heffte_input_error(_lib.last_error())
    return communicator(handle)


def comm_from_callbacks(rank, size, allgather, exchange=None):
    """
    Caller-provided transport.  allgather(mine: bytes) -> bytes of all ranks concatenated;
    exchange(sends, recvs, stream) with lists of (peer, device_ptr, nbytes) or None.
    """
    lib = _lib.load()

    def gather_cb(_context, mine, everyone, nbytes):
        try:
            data = allgather(ctypes.string_at(mine, nbytes))
            ctypes.memmove(everyone, data, len(data))
            return 0
        except Exception:
            return 1

    def exchange_cb(_context, nsend, speer, sptr, sbytes, nrecv, rpeer, rptr, rbytes, stream):
        if exchange is None:
            return 1
        try:
            sends = [(speer[i], sptr[i], sbytes[i]) for i in range(nsend)]
            recvs = [(rpeer[i], rptr[i], rbytes[i]) for i in range(nrecv)]
            exchange(sends, recvs, stream)
            return 0
        except Exception:
            return 1

    g, e = _lib.ALLGATHER_FN(gather_cb), _lib.EXCHANGE_FN(exchange_cb)
    handle = ctypes.c_void_p()
    if lib.heffte_comm_create_callbacks(rank, size, g, e, None, ctypes.byref(handle)) != 0:
        raise heffte_input_error(_lib.last_error())
    return communicator(handle, keepalive=(g, e))


# ----------------------------------------------------------------------------------------------------------------
# plans
# ----------------------------------------------------------------------------------------------------------------
def _iptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _is_torch(x):
    return type(x).__module__.startswith("torch")


_DTYPE_INFO = {  # name -> (precision, is_complex)
    "float32": (0, False), "complex64": (0, True), "float64": (1, False), "complex128": (1, True),
}


def _dtype_name(x):
    return str(x.dtype).replace("torch.", "")


def _create(backend_tag, inbox, outbox, r2c_direction, comm, options, stream):
    if backend_tag not in backend.valid:
        raise heffte_input_error("Invalid backend, this package implements heffte.backend.b200 / b200_cos / b200_sin / b200_cos1")
    lib = _lib.load()
    plan = heffte_fft_plan()
    plan.fft_comm = comm
    plan.backend_tag = backend_tag
    plan.use_r2c = r2c_direction >= 0
    plan.plan = LP_plan()
    chosen = options if options is not None else plan_options(backend_tag)
    opts = chosen.as_struct()
    herr = lib.heffte_plan_create_subcomm(backend_tag, ctypes.c_void_p(stream or 0), _iptr(inbox.low), _iptr(inbox.high), _iptr(inbox.order),
                                          _iptr(outbox.low), _iptr(outbox.high), _iptr(outbox.order), r2c_direction,
                                          comm.handle, ctypes.byref(opts), getattr(chosen, "num_subranks", -1), ctypes.byref(plan.plan))
    if herr != 0:
        plan.plan = None
        raise heffte_input_error("heFFTe encountered internal error with code: {0:1d} ({1})".format(herr, _lib.last_error()))
    return plan


def fft3d(backend_tag, inbox, outbox, comm, options=None, stream=None):
    """reference python/heffte.py:133-162: heffte.fft3d(backend_tag, inbox, outbox, comm)"""
    return _create(backend_tag, inbox, outbox, -1, comm, options, stream)


def fft3d_r2c(backend_tag, inbox, outbox, r2c_direction, comm, options=None, stream=None):
    """reference python/heffte.py:164-196"""
    if r2c_direction not in [0, 1, 2]:
        raise heffte_input_error("fft3d_r2c() called with invalid r2c_direction, must use 0, 1, or 2")
    return _create(backend_tag, inbox, outbox, r2c_direction, comm, options, stream)


class heffte_fft_plan:
    """reference python/heffte.py:198-407"""

    def __init__(self):
        self.plan = None

    def __del__(self):
        try:
            if self.plan:
                _lib.load().heffte_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    def size_inbox(self):
        return _lib.load().heffte_size_inbox64(self.plan)

    def size_outbox(self):
        return _lib.load().heffte_size_outbox64(self.plan)

    def size_workspace(self):
        return _lib.load().heffte_size_workspace64(self.plan)

    def uses_peer_memory(self, precision=1):
        return _lib.load().heffte_b200_uses_peer_memory(self.plan, precision) == 1

    def stage_timing(self, enable=True):
        _lib.load().heffte_b200_stage_timing(self.plan, int(enable))

    def stage_times(self):
        """[(name, ms, local HBM bytes, bytes sent over NVLink)] of the most recent transform (peer-memory mode)"""
        n = 32
        names = ctypes.create_string_buffer(40 * n)
        ms = (ctypes.c_double * n)()
        loc = (ctypes.c_longlong * n)()
        sent = (ctypes.c_longlong * n)()
        k = _lib.load().heffte_b200_stage_times(self.plan, n, names, ms, loc, sent)
        return [(names.raw[40 * i:40 * i + 40].split(b"\0")[0].decode(), ms[i], loc[i], sent[i]) for i in range(max(k, 0))]

    def get_scale_factor(self, scaling):
        return _lib.load().heffte_get_scale_factor(self.plan, scaling)

    def _ptr(self, x):
        if x is None:
            return ctypes.c_void_p(0)
        if _is_torch(x):
            return ctypes.c_void_p(x.data_ptr())
        return ctypes.c_void_p(x.ctypes.data)

    def _check(self, inarray, outarray, forward, batch):
        if inarray.dtype is None or _dtype_name(inarray) not in _DTYPE_INFO or _dtype_name(outarray) not in _DTYPE_INFO:
            raise heffte_input_error("use float32, float64, complex64, or complex128 arrays")
        pin, cin = _DTYPE_INFO[_dtype_name(inarray)]
        pout, cout = _DTYPE_INFO[_dtype_name(outarray)]
        if pin != pout:
            raise heffte_input_error("input and output arrays must have the same precision")
        real_transform = self.backend_tag != backend.b200
        if real_transform:
            if cin or cout:
                raise heffte_input_error("the cosine / sine transforms work with real arrays")
        elif self.use_r2c:
            if forward and (cin or not cout):
                raise heffte_input_error("forward() called with r2c can use only real inarray types and complex outarray types")
            if not forward and (not cin or cout):
                raise heffte_input_error("backward() called with r2c needs a complex inarray and a real outarray")
        else:
            if forward and not cout:
                raise heffte_input_error("forward() needs a complex outarray")
            if not forward and not cin:
                raise heffte_input_error("backward() needs a complex inarray")
        nin = self.size_inbox() if forward else self.size_outbox()
        nout = self.size_outbox() if forward else self.size_inbox()
        if _numel(inarray) != batch * nin or _numel(outarray) != batch * nout:
            raise heffte_input_error(("forward" if forward else "backward") + "() called with invalid array size")
        if _is_torch(inarray) != _is_torch(outarray):
            raise heffte_input_error("input and output must both be torch CUDA tensors or both be numpy arrays")
        # raw pointers are handed to the library: the arrays must be dense (the numpy input is copied if it is not)
        if _is_torch(inarray):
            if not inarray.is_contiguous() or not outarray.is_contiguous():
                raise heffte_input_error("torch tensors must be contiguous")
        elif not outarray.flags["C_CONTIGUOUS"]:
            raise heffte_input_error("the output numpy array must be C-contiguous")
        return pin, cin, cout

    def _check_workspace(self, workspace, outarray, forward, batch):
        """reference python/heffte.py: the workspace holds size_workspace() entries of the complex (r2r: real) type"""
        if workspace is None:
            return
        if _is_torch(workspace):
            if not workspace.is_cuda:
                raise heffte_input_error("a torch workspace must live on the GPU")
            if not workspace.is_contiguous():
                raise heffte_input_error("the workspace must be contiguous")
        elif not workspace.flags["C_CONTIGUOUS"]:
            raise heffte_input_error("the workspace must be contiguous")
        if _dtype_name(workspace) not in _DTYPE_INFO:
            raise heffte_input_error("use float32, float64, complex64, or complex128 arrays")
        wprec, wcomplex = _DTYPE_INFO[_dtype_name(workspace)]
        spectral = outarray if forward else None
        need_complex = self.backend_tag == backend.b200
        if wcomplex != need_complex:
            raise heffte_input_error("the workspace must be %s" % ("complex" if need_complex else "real"))
        if spectral is not None and wprec != _DTYPE_INFO[_dtype_name(spectral)][0]:
            raise heffte_input_error("the workspace must have the precision of the transform")
        if _numel(workspace) < batch * self.size_workspace():
            raise heffte_input_error("the workspace is smaller than size_workspace()")

    def _run(self, forward, inarray, outarray, workspace, scaling, batch):
        if scaling not in (0, 1, 2):
            raise heffte_input_error(("forward" if forward else "backward") + "() called with invalid scaling")
        precision, cin, cout = self._check(inarray, outarray, forward, batch)
        self._check_workspace(workspace, outarray, forward, batch)
        lib = _lib.load()
        direction = 0 if forward else 1
        if _is_torch(inarray):
            if not (inarray.is_cuda and outarray.is_cuda):
                raise heffte_input_error("torch tensors must live on the GPU (numpy arrays take the host path)")
            src = inarray
            if self.backend_tag == backend.b200 and not self.use_r2c and not _DTYPE_INFO[_dtype_name(inarray)][1]:
                src = inarray.to(outarray.dtype)   # real input of a c2c plan: zero imaginary part (reference heffte_backend_cuda.h:527-536)
            dst = outarray
            tmp = None
            if self.backend_tag == backend.b200 and not self.use_r2c and not cout:
                import torch
                tmp = torch.empty(outarray.numel(), dtype=inarray.dtype, device=outarray.device)
                dst = tmp
            rc = lib.heffte_execute(self.plan, precision, direction, batch, self._ptr(src), self._ptr(dst), self._ptr(workspace), scaling)
            if rc == 0 and tmp is not None:
                outarray.copy_(tmp.real.reshape(outarray.shape))
        else:
            src = np.ascontiguousarray(inarray)
            if self.backend_tag == backend.b200 and not self.use_r2c and not cin:
                src = src.astype(outarray.dtype)
            dst = outarray
            if self.backend_tag == backend.b200 and not self.use_r2c and not cout:
                dst = np.empty(outarray.size, dtype=inarray.dtype)
            rc = lib.heffte_execute_host(self.plan, precision, direction, batch, self._ptr(src), self._ptr(dst), scaling)
            if rc == 0 and dst is not outarray:
                outarray[...] = dst.real.reshape(outarray.shape)
        if rc != 0:
            raise heffte_input_error("heFFTe(b200) transform failed with code %d: %s" % (rc, _lib.last_error()))

    def convolve(self, inarray, outarray, multiplier=None, scaling=scale.full):
        """
        Fused spectral operator of a complex-to-complex plan: outarray = backward(forward(inarray) * factor(scaling) * M), M = the
        spectrum itself (multiplier None: the x[i] *= x[i] of the reference's benchmarks/convolution.cpp:86-97) or a device array
        laid out over convolve_box().  One plan-level call: the reshapes around the product are not executed.
        """
        self.convolve_buffered(inarray, outarray, None, multiplier, scaling)

    def convolve_buffered(self, inarray, outarray, workspace, multiplier=None, scaling=scale.full):
        if self.backend_tag != backend.b200 or self.use_r2c:
            raise heffte_input_error("convolve() is defined for complex-to-complex plans")
        if not (_is_torch(inarray) and _is_torch(outarray)) and not (not _is_torch(inarray) and not _is_torch(outarray)):
            raise heffte_input_error("input and output must both be torch CUDA tensors or both be numpy arrays")
        if _dtype_name(inarray) not in _DTYPE_INFO or _dtype_name(inarray) != _dtype_name(outarray) or not _DTYPE_INFO[_dtype_name(inarray)][1]:
            raise heffte_input_error("convolve() works on complex64 or complex128 arrays of one type")
        if _numel(inarray) != self.size_inbox() or _numel(outarray) != self.size_inbox():
            raise heffte_input_error("convolve() called with invalid array size")
        if _is_torch(inarray) and not (inarray.is_contiguous() and outarray.is_contiguous()):
            raise heffte_input_error("torch tensors must be contiguous")
        if multiplier is not None:
            lo, hi, _ = self.convolve_box()
            count = max(0, hi[0] - lo[0] + 1) * max(0, hi[1] - lo[1] + 1) * max(0, hi[2] - lo[2] + 1)
            if _numel(multiplier) != count or _dtype_name(multiplier) != _dtype_name(inarray):
                raise heffte_input_error("the multiplier covers convolve_box() with the type of the data")
        precision = _DTYPE_INFO[_dtype_name(inarray)][0]
        rc = _lib.load().heffte_convolve(self.plan, precision, self._ptr(inarray), self._ptr(outarray), self._ptr(workspace), self._ptr(multiplier), scaling)
        if rc != 0:
            raise heffte_input_error("heFFTe(b200) convolve failed with code %d: %s" % (rc, _lib.last_error()))

    def convolve_box(self):
        """(low, high, order) of this rank's part of the spectrum while convolve() applies the multiplier"""
        low, high, order = (ctypes.c_longlong * 3)(), (ctypes.c_longlong * 3)(), (ctypes.c_int * 3)()
        _lib.load().heffte_convolve_box(self.plan, low, high, order)
        return [int(v) for v in low], [int(v) for v in high], [int(v) for v in order]

    def prepare(self, precision=1, batch=1):
        """collective: set up the peer-memory data plane for transforms of up to `batch` entries ahead of the first transform"""
        rc = _lib.load().heffte_b200_prepare(self.plan, precision, batch)
        if rc != 0:
            raise heffte_input_error("heFFTe(b200) prepare failed with code %d: %s" % (rc, _lib.last_error()))

    def register_buffer(self, array):
        """collective: register a device array this plan will be asked to write (every rank its own); the other GPUs then store
        their part of a result straight into it.  Every rank must pass its registered array as the output of the same call.
        Returns True when the array was registered, False when the plan cannot share it (nothing changes then)."""
        if _dtype_name(array) not in _DTYPE_INFO:
            raise heffte_input_error("use float32, float64, complex64, or complex128 arrays")
        precision, _ = _DTYPE_INFO[_dtype_name(array)]
        if _is_torch(array):
            if not array.is_contiguous():
                raise heffte_input_error("register_buffer needs a contiguous array")
            nbytes = array.numel() * array.element_size()
        else:
            nbytes = array.nbytes
        rc = _lib.load().heffte_b200_register_buffer(self.plan, precision, self._ptr(array), nbytes)
        if rc not in (0, 2):
            raise heffte_input_error("heFFTe(b200) register_buffer failed with code %d: %s" % (rc, _lib.last_error()))
        return rc == 0

    def unregister_buffer(self, array):
        """local: forget a registered array (call it before the array is released)"""
        precision, _ = _DTYPE_INFO[_dtype_name(array)]
        _lib.load().heffte_b200_unregister_buffer(self.plan, precision, self._ptr(array))

    def forward(self, inarray, outarray, scaling=scale.none, batch=1):
        self._run(True, inarray, outarray, None, scaling, batch)

    def forward_buffered(self, inarray, outarray, workspace, scaling=scale.none, batch=1):
        self._run(True, inarray, outarray, workspace, scaling, batch)

    def backward(self, inarray, outarray, scaling=scale.none, batch=1):
        self._run(False, inarray, outarray, None, scaling, batch)

    def backward_buffered(self, inarray, outarray, workspace, scaling=scale.none, batch=1):
        self._run(False, inarray, outarray, workspace, scaling, batch)


def _numel(x):
    return x.numel() if _is_torch(x) else x.size


# ----------------------------------------------------------------------------------------------------------------
# geometry helpers used by the benchmark (reference include/heffte_geometry.h:409-436, 643-691)
# ----------------------------------------------------------------------------------------------------------------
def proc_setup_min_surface(world, nprocs):
    lib = _lib.load()
    w = np.array(world.nine(), dtype=np.int32)
    g = np.zeros(3, dtype=np.int32)
    lib.heffte_b200_proc_setup_min_surface(_iptr(w), nprocs, _iptr(g))
    return [int(v) for v in g]


def split_world(world, grid):
    lib = _lib.load()
    n = int(grid[0]) * int(grid[1]) * int(grid[2])
    w = np.array(world.nine(), dtype=np.int32)
    g = np.array(grid, dtype=np.int32)
    out = np.zeros(9 * n, dtype=np.int32)
    lib.heffte_b200_split_world(_iptr(w), _iptr(g), _iptr(out))
    return [box3d(out[9 * i:9 * i + 3], out[9 * i + 3:9 * i + 6], out[9 * i + 6:9 * i + 9]) for i in range(n)]


def make_procgrid(nprocs):
    lib = _lib.load()
    g = np.zeros(2, dtype=np.int32)
    lib.heffte_b200_make_procgrid(nprocs, _iptr(g))
    return [int(v) for v in g]


def logic_plan(inboxes, outboxes, r2c_direction=-1, use_reorder=False, algorithm=0, use_pencils=True, subranks=-1, rank=0):
    """Pure host planning: returns (shapes[8][nranks][9], fft_direction, index_count), see include/heffte_b200.h."""
    lib = _lib.load()
    n = len(inboxes)
    ib = np.array([b.nine() for b in inboxes], dtype=np.int32).reshape(-1)
    ob = np.array([b.nine() for b in outboxes], dtype=np.int32).reshape(-1)
    shapes = np.zeros(8 * n * 9, dtype=np.int32)
    fdir = np.zeros(3, dtype=np.int32)
    count = ctypes.c_longlong(0)
    rc = lib.heffte_b200_logic_plan(n, _iptr(ib), _iptr(ob), r2c_direction, int(use_reorder), algorithm, int(use_pencils), subranks, rank,
                                    _iptr(shapes), _iptr(fdir), ctypes.byref(count))
    if rc != 0:
        raise heffte_input_error(_lib.last_error())
    return shapes.reshape(8, n, 9).tolist(), fdir.tolist(), count.value


def execution_plan(inboxes, outboxes, r2c_direction=-1, use_reorder=False, algorithm=0, use_pencils=None, subranks=-1, rank=0):
    """Pure host planning: the plan a b200 transform executes (no reorder, traffic-balanced): (shapes[8][nranks][9], fft_direction, swaps)."""
    lib = _lib.load()
    n = len(inboxes)
    ib = np.array([b.nine() for b in inboxes], dtype=np.int32).reshape(-1)
    ob = np.array([b.nine() for b in outboxes], dtype=np.int32).reshape(-1)
    shapes = np.zeros(8 * n * 9, dtype=np.int32)
    fdir = np.zeros(3, dtype=np.int32)
    swaps = np.zeros(1, dtype=np.int32)
    rc = lib.heffte_b200_execution_plan(n, _iptr(ib), _iptr(ob), r2c_direction, int(use_reorder), algorithm, 2 if use_pencils is None else int(use_pencils), subranks, rank,
                                        _iptr(shapes), _iptr(fdir), _iptr(swaps))
    if rc != 0:
        raise heffte_input_error(_lib.last_error())
    return shapes.reshape(8, n, 9).tolist(), fdir.tolist(), int(swaps[0])


def reshape_pieces(inboxes, outboxes, me, receive):
    lib = _lib.load()
    n = len(inboxes)
    ib = np.array([b.nine() for b in inboxes], dtype=np.int32).reshape(-1)
    ob = np.array([b.nine() for b in outboxes], dtype=np.int32).reshape(-1)
    out = np.zeros(14 * (n + 1), dtype=np.int64)
    k = lib.heffte_b200_reshape_pieces(n, _iptr(ib), _iptr(ob), me, int(receive), out.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), n + 1)
    if k < 0:
        raise heffte_input_error(_lib.last_error())
    keys = ["peer", "offset", "size0", "size1", "size2", "line", "plane", "buff_line", "buff_plane", "map0", "map1", "map2", "count", "buffer_offset"]
    return [dict(zip(keys, [int(v) for v in out[14 * i:14 * i + 14]])) for i in range(k)]


def plan_sizes(kind, inboxes, outboxes, rank, r2c_direction=-1, use_reorder=False, algorithm=0, use_pencils=True, subranks=-1):
    """(size_inbox, size_outbox, size_workspace) of the plan rank `rank` would build, no device needed."""
    lib = _lib.load()
    n = len(inboxes)
    ib = np.array([b.nine() for b in inboxes], dtype=np.int32).reshape(-1)
    ob = np.array([b.nine() for b in outboxes], dtype=np.int32).reshape(-1)
    a, b, c = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_longlong(0)
    rc = lib.heffte_b200_plan_sizes(kind, n, _iptr(ib), _iptr(ob), r2c_direction, int(use_reorder), algorithm, int(use_pencils), subranks, rank,
                                    ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    if rc != 0:
        raise heffte_input_error(_lib.last_error())
    return a.value, b.value, c.value
