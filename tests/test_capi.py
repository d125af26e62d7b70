"""
The drop-in boundary on a host without a GPU: libheffte_b200.so loads, exports every entry point the headers in
include/ declare, keeps the struct layouts / constants of the reference C interface (include/heffte_c_defines.h:54-162),
follows the reference's return-code conventions (src/heffte_c.cpp:232, 273-277, 337) and fails loudly -- never falls
back to a CPU path -- when a compute call is made without a CUDA device.
"""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import heffte_oracle as O
from tests.helpers import to_h

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text):
        name = m.group(1)
        if name.startswith(("heffte_", "b200_")) and not name.endswith(("_fn",)):
            names.add(name)
    return names


@pytest.mark.parametrize("header", ["heffte_b200.h", "heffte_b200_kernels.h"])
def test_every_declared_symbol_is_exported(lib, header):
    names = _declared_functions(header)
    assert len(names) > 10
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_and_constants(lib):
    from heffte_b200 import _lib, heffte as H
    assert ctypes.sizeof(_lib.heffte_plan_options) == 16            # four ints (heffte_c_defines.h:113-125)
    assert [f[0] for f in _lib.heffte_plan_options._fields_] == ["use_reorder", "algorithm", "use_pencils", "use_gpu_aware"]
    assert [f[0] for f in _lib.heffte_fft_plan_struct._fields_] == ["backend_type", "using_r2c", "fft"]
    assert (H.backend.stock, H.backend.fftw, H.backend.mkl, H.backend.cufft, H.backend.rocm) == (0, 1, 2, 10, 11)
    assert (H.scale.none, H.scale.full, H.scale.symmetric) == (0, 1, 2)
    assert (H.reshape_algorithm.alltoallv, H.reshape_algorithm.p2p_plined, H.reshape_algorithm.p2p, H.reshape_algorithm.alltoall) == (0, 1, 2, 3)
    opts = _lib.heffte_plan_options()
    assert lib.heffte_set_default_options(H.backend.b200, ctypes.byref(opts)) == 0
    assert (opts.use_reorder, opts.algorithm, opts.use_pencils, opts.use_gpu_aware) == (0, 0, 1, 1)   # cufft defaults (heffte_backend_cuda.h:854-857)
    assert lib.heffte_set_default_options(H.backend.b200_cos, ctypes.byref(opts)) == 0 and opts.use_reorder == 1
    assert lib.heffte_set_default_options(H.backend.fftw, ctypes.byref(opts)) == 1                     # backend not in this library


def test_plan_lifecycle_and_return_codes(lib):
    from heffte_b200 import _lib, heffte as H
    world = to_h(O.world_box((4, 4, 4)))
    comm = H.comm_self()
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    plan = _lib.LP_plan()
    # invalid backend -> 1 (src/heffte_c.cpp:232)
    assert lib.heffte_plan_create(H.backend.cufft, ip(world.low), ip(world.high), None, ip(world.low), ip(world.high), None, comm.handle, None, ctypes.byref(plan)) == 1
    # null communicator / bad geometry -> 2 (exception path, src/heffte_c.cpp:273-277)
    assert lib.heffte_plan_create(H.backend.b200, ip(world.low), ip(world.high), None, ip(world.low), ip(world.high), None, None, None, ctypes.byref(plan)) == 2
    other = to_h(O.world_box((4, 4, 5)))
    assert lib.heffte_plan_create(H.backend.b200, ip(world.low), ip(world.high), None, ip(other.low), ip(other.high), None, comm.handle, None, ctypes.byref(plan)) == 2
    assert "box" in _lib.last_error().lower() or "world" in _lib.last_error().lower()
    # a valid plan: NULL order and NULL options are accepted (include/heffte_c.h:67)
    assert lib.heffte_plan_create(H.backend.b200, ip(world.low), ip(world.high), None, ip(world.low), ip(world.high), None, comm.handle, None, ctypes.byref(plan)) == 0
    assert lib.heffte_size_inbox(plan) == 64 and lib.heffte_size_outbox(plan) == 64 and lib.heffte_size_workspace(plan) >= 64
    assert lib.heffte_get_backend(plan) == H.backend.b200 and lib.heffte_is_r2c(plan) == 0
    assert abs(lib.heffte_get_scale_factor(plan, 1) - 1.0 / 64) < 1e-16 and abs(lib.heffte_get_scale_factor(plan, 2) - 0.125) < 1e-16
    # corrupt handle -> 3 (src/heffte_c.cpp:337)
    saved = plan.contents.backend_type
    plan.contents.backend_type = 77
    assert lib.heffte_plan_destroy(plan) == 3
    plan.contents.backend_type = saved
    assert lib.heffte_plan_destroy(plan) == 0
    # r2c plan and its direction check
    cworld = to_h(O.world_box((4, 4, 4)).r2c(1))
    assert lib.heffte_plan_create_r2c(H.backend.b200, ip(world.low), ip(world.high), None, ip(cworld.low), ip(cworld.high), None, 1, comm.handle, None, ctypes.byref(plan)) == 0
    assert lib.heffte_is_r2c(plan) == 1 and lib.heffte_size_outbox(plan) == 48
    assert lib.heffte_plan_destroy(plan) == 0
    assert lib.heffte_plan_create_r2c(H.backend.b200, ip(world.low), ip(world.high), None, ip(cworld.low), ip(cworld.high), None, 7, comm.handle, None, ctypes.byref(plan)) == 2


def test_thread_ranks_need_their_own_stream(lib):
    """ranks that are host threads of one process share the default stream: plan creation on it is refused (code 2, the
    exception path of src/heffte_c.cpp:273-277) before any collective call is made"""
    from heffte_b200 import _lib, heffte as H
    comms = H.comm_threads(2)
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    low, high = np.array([0, 0, 0], dtype=np.int32), np.array([3, 3, 1], dtype=np.int32)
    plan = _lib.LP_plan()
    rc = lib.heffte_plan_create(H.backend.b200, ip(low), ip(high), None, ip(low), ip(high), None, comms[0].handle, None, ctypes.byref(plan))
    assert rc == 2 and "stream" in _lib.last_error()


def test_real_side_entry_points_and_helpers_are_exported(lib):
    """entry points added next to the reference's: streams, sub-box copy, ranged 1-D execution, executed-plan introspection"""
    for name in ("b200_stream_create", "b200_stream_destroy", "b200_copy_subbox", "b200_fft1d_execute_range", "heffte_b200_execution_plan",
                 "heffte_forward_d2z", "heffte_backward_z2d", "heffte_forward_s2c", "heffte_backward_c2s"):
        assert hasattr(lib, name), name


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry point must fail loudly (never compute on the host)."""
    if lib.b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from heffte_b200 import _lib, heffte as H
    d = _lib.b200_fft1d_desc(1, 0, 16, 4, 1, _lib.b200_line_geom(1, 16, 0), _lib.b200_line_geom(1, 16, 0))
    plan = ctypes.c_void_p()
    assert lib.b200_fft1d_create(ctypes.byref(d), ctypes.byref(plan)) == 5        # B200_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.b200_last_error()
    world = to_h(O.world_box((4, 4, 4)))
    fft = H.fft3d(H.backend.b200, world, world, H.comm_self())
    x = np.zeros(64, dtype=np.complex128)
    with pytest.raises(H.heffte_input_error):
        fft.forward(x, x.copy())
    buf = np.zeros(64)
    assert lib.b200_scale(1, 64, buf.ctypes.data, 2.0, None) != 0


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under heffte_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "heffte_b200")
    for base, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, name), errors="ignore").read()
                assert "oracle" not in text.replace("oracle/ is", ""), os.path.join(base, name)
