"""ctypes loader of libheffte_b200.so.  There is no fallback: if the CUDA library is missing the import fails loudly."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libheffte_b200.so")


class heffte_plan_options(ctypes.Structure):
    # reference include/heffte_c_defines.h:113-125
    _fields_ = [("use_reorder", ctypes.c_int), ("algorithm", ctypes.c_int), ("use_pencils", ctypes.c_int), ("use_gpu_aware", ctypes.c_int)]


class heffte_fft_plan_struct(ctypes.Structure):
    # reference include/heffte_c_defines.h:133-146
    _fields_ = [("backend_type", ctypes.c_int), ("using_r2c", ctypes.c_int), ("fft", ctypes.c_void_p)]


class b200_line_geom(ctypes.Structure):
    _fields_ = [("stride", ctypes.c_longlong), ("stride_a", ctypes.c_longlong), ("stride_b", ctypes.c_longlong)]


class b200_fft1d_desc(ctypes.Structure):
    _fields_ = [("precision", ctypes.c_int), ("kind", ctypes.c_int), ("n", ctypes.c_longlong),
                ("count_a", ctypes.c_longlong), ("count_b", ctypes.c_longlong), ("in_", b200_line_geom), ("out", b200_line_geom)]


LP_plan = ctypes.POINTER(heffte_fft_plan_struct)
ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)
EXCHANGE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_void_p),
                               ctypes.POINTER(ctypes.c_size_t), ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_void_p),
                               ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p)

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("heffte_b200: %s is missing; build it with `python -m heffte_b200.build` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    c_int, c_ll, c_vp, c_dbl = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_double
    ip = ctypes.POINTER(c_int)

    def sig(name, restype, *argtypes):
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = list(argtypes)

    sig("b200_last_error", ctypes.c_char_p)
    sig("heffte_last_error", ctypes.c_char_p)
    sig("b200_launch_count", c_ll)
    sig("b200_device_count", c_int)
    sig("b200_device_alloc", c_int, ctypes.c_size_t, ctypes.POINTER(c_vp))
    sig("b200_device_free", c_int, c_vp)
    sig("b200_copy_to_device", c_int, c_vp, c_vp, ctypes.c_size_t, c_vp)
    sig("b200_copy_to_host", c_int, c_vp, c_vp, ctypes.c_size_t, c_vp)
    sig("b200_copy_on_device", c_int, c_vp, c_vp, ctypes.c_size_t, c_vp)
    sig("b200_copy_any", c_int, c_vp, c_vp, ctypes.c_size_t)
    sig("b200_stream_synchronize", c_int, c_vp)
    sig("b200_device_set", c_int, c_int)
    sig("b200_fft1d_create", c_int, ctypes.POINTER(b200_fft1d_desc), ctypes.POINTER(c_vp))
    sig("b200_fft1d_destroy", c_int, c_vp)
    sig("b200_fft1d_execute", c_int, c_vp, c_int, c_vp, c_vp, c_dbl, c_vp)
    sig("b200_fft1d_execute_range", c_int, c_vp, c_int, c_vp, c_vp, c_dbl, c_vp, c_ll, c_ll)
    sig("b200_fft1d_kernel_name", ctypes.c_char_p, c_vp)
    sig("b200_direct_pack", c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_ll, c_vp, c_vp, c_vp)
    sig("b200_direct_unpack", c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_ll, c_vp, c_vp, c_vp)
    sig("b200_transpose_unpack", c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_ll, c_ll, c_ll, c_int, c_int, c_int, c_vp, c_vp, c_vp)
    sig("b200_scale", c_int, c_int, c_ll, c_vp, c_dbl, c_vp)
    sig("b200_convert_r2c", c_int, c_int, c_ll, c_vp, c_vp, c_vp)
    sig("b200_convert_c2r", c_int, c_int, c_ll, c_vp, c_vp, c_vp)

    sig("heffte_comm_create_self", c_int, ctypes.POINTER(c_vp))
    sig("heffte_comm_nccl_unique_id", c_int, c_vp)
    sig("heffte_comm_create_nccl", c_int, c_int, c_int, c_vp, ctypes.POINTER(c_vp))
    sig("heffte_comm_create_callbacks", c_int, c_int, c_int, ALLGATHER_FN, EXCHANGE_FN, c_vp, ctypes.POINTER(c_vp))
    sig("heffte_comm_create_threads", c_int, c_int, ip, ctypes.POINTER(c_vp))
    sig("heffte_comm_rank", c_int, c_vp)
    sig("heffte_comm_size", c_int, c_vp)
    sig("heffte_comm_destroy", c_int, c_vp)

    sig("heffte_set_default_options", c_int, c_int, ctypes.POINTER(heffte_plan_options))
    sig("heffte_plan_create", c_int, c_int, ip, ip, ip, ip, ip, ip, c_vp, ctypes.POINTER(heffte_plan_options), ctypes.POINTER(LP_plan))
    sig("heffte_plan_create_r2c", c_int, c_int, ip, ip, ip, ip, ip, ip, c_int, c_vp, ctypes.POINTER(heffte_plan_options), ctypes.POINTER(LP_plan))
    sig("heffte_plan_create_stream", c_int, c_int, c_vp, ip, ip, ip, ip, ip, ip, c_int, c_vp, ctypes.POINTER(heffte_plan_options), ctypes.POINTER(LP_plan))
    sig("heffte_plan_create_subcomm", c_int, c_int, c_vp, ip, ip, ip, ip, ip, ip, c_int, c_vp, ctypes.POINTER(heffte_plan_options), c_int,
        ctypes.POINTER(LP_plan))
    sig("heffte_plan_destroy", c_int, LP_plan)
    for name in ("heffte_size_inbox", "heffte_size_outbox", "heffte_size_workspace", "heffte_get_backend", "heffte_is_r2c"):
        sig(name, c_int, LP_plan)
    for name in ("heffte_size_inbox64", "heffte_size_outbox64", "heffte_size_workspace64"):
        sig(name, c_ll, LP_plan)
    sig("heffte_get_scale_factor", c_dbl, LP_plan, c_int)
    for name in ("s2c", "c2c", "d2z", "z2z"):
        sig("heffte_forward_" + name, None, LP_plan, c_vp, c_vp, c_int)
        sig("heffte_forward_" + name + "_buffered", None, LP_plan, c_vp, c_vp, c_vp, c_int)
    for name in ("c2s", "c2c", "z2d", "z2z"):
        sig("heffte_backward_" + name, None, LP_plan, c_vp, c_vp, c_int)
        sig("heffte_backward_" + name + "_buffered", None, LP_plan, c_vp, c_vp, c_vp, c_int)
    for name in ("forward_s2s", "forward_d2d", "backward_s2s", "backward_d2d"):
        sig("heffte_" + name + "_buffered", None, LP_plan, c_vp, c_vp, c_vp, c_int)
    sig("heffte_b200_uses_peer_memory", c_int, LP_plan, c_int)
    sig("heffte_b200_stage_timing", c_int, LP_plan, c_int)
    sig("heffte_b200_stage_times", c_int, LP_plan, c_int, ctypes.c_char_p, ctypes.POINTER(c_dbl), ctypes.POINTER(c_ll), ctypes.POINTER(c_ll))
    sig("b200_fft1d_execute_scatter", c_int, c_vp, c_int, c_vp, c_vp, c_dbl, c_vp)
    sig("b200_scatter_copy", c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_ll, c_vp, c_vp, c_vp)
    sig("b200_peer_barrier", c_int, c_int, c_int, ctypes.POINTER(c_vp), c_vp, ctypes.c_ulonglong, c_vp)
    sig("heffte_execute", c_int, LP_plan, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int)
    sig("heffte_execute_host", c_int, LP_plan, c_int, c_int, c_int, c_vp, c_vp, c_int)
    sig("heffte_convolve", c_int, LP_plan, c_int, c_vp, c_vp, c_vp, c_vp, c_int)
    sig("heffte_convolve_box", c_int, LP_plan, ctypes.POINTER(c_ll), ctypes.POINTER(c_ll), ip)
    sig("heffte_b200_prepare", c_int, LP_plan, c_int, c_int)
    sig("heffte_b200_register_buffer", c_int, LP_plan, c_int, c_vp, ctypes.c_size_t)
    sig("heffte_b200_unregister_buffer", c_int, LP_plan, c_int, c_vp)
    sig("heffte_plan_create64", c_int, c_int, c_vp, ctypes.POINTER(c_ll), ctypes.POINTER(c_ll), ip, ctypes.POINTER(c_ll), ctypes.POINTER(c_ll), ip,
        c_int, c_vp, ctypes.POINTER(heffte_plan_options), c_int, ctypes.POINTER(LP_plan))
    sig("b200_fft1d_execute_batch", c_int, c_vp, c_int, c_vp, c_vp, c_dbl, c_vp, c_int, c_ll, c_ll)
    sig("b200_fft1d_execute_scatter_batch", c_int, c_vp, c_int, c_vp, c_vp, c_dbl, c_vp, c_int, c_ll, c_ll, c_ll, c_ll)
    sig("b200_fft1d_pairable", c_int, c_vp, c_vp)
    sig("b200_fft1d_execute_pair", c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_dbl, c_vp, c_int, c_vp, c_int, c_ll, c_ll, c_ll, c_ll, c_ll)
    sig("b200_fft1d_convolvable", c_int, c_vp)
    sig("b200_fft1d_execute_convolve", c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl, c_vp, c_int, c_ll, c_ll, c_ll, c_ll, c_ll)
    sig("b200_pointwise_multiply", c_int, c_int, c_ll, c_vp, c_vp, c_dbl, c_vp)
    sig("b200_scatter_copy_batch", c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_ll, c_vp, c_vp, c_vp, c_int, c_ll, c_ll, c_ll, c_ll)
    sig("b200_copy_subboxes", c_int, c_int, c_int, ctypes.POINTER(c_ll), ctypes.POINTER(c_ll), ctypes.POINTER(c_ll), ctypes.POINTER(c_ll),
        c_ll, c_ll, c_vp, c_vp, c_vp, c_int, c_ll, c_ll)
    sig("b200_peer_timed_out", ctypes.c_ulonglong)

    sig("heffte_b200_logic_plan", c_int, c_int, ip, ip, c_int, c_int, c_int, c_int, c_int, c_int, ip, ip, ctypes.POINTER(c_ll))
    sig("heffte_b200_make_procgrid", None, c_int, ip)
    sig("heffte_b200_proc_setup_min_surface", None, ip, c_int, ip)
    sig("heffte_b200_split_world", None, ip, ip, ip)
    sig("heffte_b200_execution_plan", c_int, c_int, ip, ip, c_int, c_int, c_int, c_int, c_int, c_int, ip, ip, ip)
    sig("heffte_b200_reshape_pieces", c_int, c_int, ip, ip, c_int, c_int, ctypes.POINTER(c_ll), c_int)
    sig("heffte_b200_plan_sizes", c_int, c_int, c_int, ip, ip, c_int, c_int, c_int, c_int, c_int, c_int,
        ctypes.POINTER(c_ll), ctypes.POINTER(c_ll), ctypes.POINTER(c_ll))
    _lib = lib
    return lib


def last_error():
    return load().heffte_last_error().decode()
