"""ctypes binding of ``libtransflow_b200.so`` (the C ABI declared in ``include/transflow_b200.h``).

There is no CPU fallback: importing this module without the built library raises, and every
call on a machine without an sm_100 GPU raises ``RuntimeError`` (TF_ERR_UNSUPPORTED_ARCH /
TF_ERR_CUDA).  PyTorch tensors are only the owners of device memory; the library receives raw
device pointers plus the current CUDA stream.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TFB200_LIB") or os.path.join(_HERE, "libtransflow_b200.so")   # (override: A/B of two builds)

TF_OK = 0
TF_ERR_INVALID_ARG = -1
TF_ERR_CUDA = -2
TF_ERR_UNSUPPORTED_ARCH = -3
TF_ERR_SHAPE = -4
TF_ERR_INDEX = -5

LAYER_KINDS = {"moveref": 0, "sum": 1, "static": 2, "introduction": 3}
RESET_MODES = {"off": 0, "random": 1, "constant": 2, "linear": 3}


class TransflowLibraryMissing(ImportError):
    pass


class LayerConfigStruct(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("transparent_pixels_can_move", C.c_int32),
        ("pixels_can_move_to_empty_spot", C.c_int32),
        ("pixels_can_move_to_filled_spot", C.c_int32),
        ("moving_pixels_leave_empty_spot", C.c_int32),
        ("reset_mode", C.c_int32),
        ("reset_source", C.c_int32),
        ("introduce_pixels_on_empty_spots", C.c_int32),
        ("introduce_pixels_on_filled_spots", C.c_int32),
        ("introduce_moving_pixels", C.c_int32),
        ("introduce_unmoving_pixels", C.c_int32),
        ("introduce_once", C.c_int32),
        ("introduce_on_all_filled_spots", C.c_int32),
        ("introduce_on_all_empty_spots", C.c_int32),
        ("reset_constant_step", C.c_float),
        ("reset_random_factor", C.c_float),
        ("reset_linear_factor", C.c_double),
    ]


class FlowOpStruct(C.Structure):
    """``tf_flow_op``: kind (0 scale, 1 threshold, 2 clip), strong (NumPy float64 scalar), value."""
    _fields_ = [("kind", C.c_int), ("strong", C.c_int), ("value", C.c_double)]


class PixmapStruct(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("channels", C.c_int32), ("frame_number", C.c_int32)]


_vp, _i, _d, _u64, _u32 = C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint32
_pi = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every function of include/transflow_b200.h
SIGNATURES = {
    "tf_version": (_i, []),
    "tf_last_error": (C.c_char_p, []),
    "tf_device_check": (_i, [_i, _pi]),
    "tf_launch_count": (_u64, []),
    "tf_timer_enable": (_i, [_i]),
    "tf_timer_read": (_i, [_i, C.POINTER(_d), C.POINTER(_u64)]),
    "tf_gray_from_bgr": (_i, [_vp, _vp, _i, _i, _vp]),
    "tf_resize_nearest_bgr": (_i, [_vp, _i, _i, _vp, _i, _i, _vp]),
    "tf_farneback_create": (_i, [C.POINTER(_vp), _i, _i, _d, _i, _i, _i, _i, _d, _i, _i]),
    "tf_farneback_destroy": (_i, [_vp]),
    "tf_farneback_prepare": (_i, [_vp, _i, _vp, _vp]),
    "tf_farneback_solve": (_i, [_vp, _i, _i, _vp, _i, _i, _vp]),
    "tf_farneback_step": (_i, [_vp, _i, _vp, _i, _i, _vp, _i, _i, _vp]),
    "tf_farneback_step_lane": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp]),
    "tf_farneback_run": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "tf_farneback_num_levels": (_i, [_vp]),
    "tf_farneback_level_size": (_i, [_vp, _i, _pi, _pi]),
    "tf_farneback_debug_read": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "tf_farneback_algorithmic_bytes": (_d, [_vp, _i]),
    "tf_farneback_tune": (_i, [_i, _i]),
    "tf_farneback_set_debug": (_i, [_vp, _i]),
    "tf_farneback_reserve": (_i, [_vp, _i, _i]),
    "tf_floatmap_accumulate": (_i, [_vp, _vp, _vp, _i, _i, C.c_float, C.c_float, _i, _i, _vp]),
    "tf_floatmap_remap": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _u32, _i, _i, _vp]),
    "tf_hs_create": (_i, [C.POINTER(_vp), _i, _i]),
    "tf_hs_destroy": (_i, [_vp]),
    "tf_hs_run": (_i, [_vp, _vp, _vp, _vp, _d, _i, _d, _d, _vp, _i, _pi, _vp]),
    "tf_lk_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "tf_lk_destroy": (_i, [_vp]),
    "tf_lk_run": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "tf_flow_postprocess": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp]),
    "tf_flow_postprocess_to": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "tf_flow_postprocess_ex": (_i, [_vp, C.POINTER(FlowOpStruct), _i, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "tf_flow_forward_claims": (_i, [_vp, C.POINTER(FlowOpStruct), _i, _vp, _vp, _i, _i, _vp]),
    "tf_flow_from_claims": (_i, [_vp, _vp, _i, _i, _vp]),
    "tf_flow_filters": (_i, [_vp, C.POINTER(FlowOpStruct), _i, _vp, _vp, _i, _i, _vp]),
    "tf_flow_convolve": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _vp]),
    "tf_flow_merge": (_i, [C.POINTER(_vp), _i, _i, _vp, _i, _i, _vp]),
    "tf_flow_upscale": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "tf_flow_render": (_i, [_vp, _i, C.c_float, _vp, _i, _i, _vp, _i, _i, _vp]),
    "tf_device_malloc": (_i, [C.c_size_t, C.POINTER(_vp)]),
    "tf_device_free": (_i, [_vp]),
    "tf_layer_create": (_i, [C.POINTER(_vp), _i, _i, C.POINTER(LayerConfigStruct)]),
    "tf_layer_destroy": (_i, [_vp]),
    "tf_layer_set_masks": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "tf_layer_set_sources": (_i, [_vp, _i, C.POINTER(_vp), _vp]),
    "tf_layer_update": (_i, [_vp, _vp, C.POINTER(PixmapStruct), _i, _vp, _u64, _vp, _i, _u32, _vp]),
    "tf_layer_takes_claims": (_i, [_vp, _i]),
    "tf_layer_update_claims": (_i, [_vp, _vp, C.POINTER(PixmapStruct), _i, _u64, _vp, _u32, _vp]),
    "tf_layer_render": (_i, [_vp, _vp, _vp]),
    "tf_composite": (_i, [C.POINTER(_vp), _i, _u32, _vp, _i, _i, _vp]),
    "tf_layer_depth": (_i, [_vp]),
    "tf_layer_get_state": (_i, [_vp, _vp, _vp, _vp]),
    "tf_layer_set_state": (_i, [_vp, _vp, _vp, _vp]),
    "tf_layer_poll_error": (_i, [_vp, _vp]),
    "tf_layer_get_counters": (_i, [_vp, C.POINTER(_u64), _pi]),
    "tf_layer_set_counters": (_i, [_vp, _u64, _i]),
    "tf_ipc_get_handle": (_i, [_vp, _vp]),
    "tf_ipc_open_handle": (_i, [_vp, C.POINTER(_vp)]),
    "tf_ipc_close_handle": (_i, [_vp]),
    "tf_flag_signal": (_i, [_vp, _u32, _vp]),
    "tf_flag_wait_geq": (_i, [_vp, _u32, _vp]),
    "tf_copy_to_peer": (_i, [_vp, _vp, C.c_size_t, _vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise TransflowLibraryMissing(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C transflow_b200/csrc`).  transflow_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().tf_last_error().decode("utf8", "replace")


def check(status: int):
    """Map a tf_status to the Python exception the reference would have raised."""
    if status == TF_OK:
        return
    msg = last_error()
    if status in (TF_ERR_INVALID_ARG, TF_ERR_SHAPE):
        raise ValueError(msg)
    if status == TF_ERR_INDEX:
        raise IndexError(msg)
    raise RuntimeError(msg)


_raw_stream = None


def stream_ptr():
    """The current torch stream of the current device as a ``cudaStream_t`` (every call of the library takes one).
    Read through torch's raw accessor: ``torch.cuda.current_stream()`` builds a Python Stream object on every call, and
    a frame makes a dozen calls."""
    global _raw_stream
    if _raw_stream is None:
        import torch
        get_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        get_device = getattr(torch._C, "_cuda_getDevice", None)
        if get_stream is not None and get_device is not None:
            _raw_stream = lambda: get_stream(get_device())          # noqa: E731
        else:
            _raw_stream = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    return C.c_void_p(_raw_stream())


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, int):          # a raw device address (e.g. a CUDA-IPC mapping of peer memory)
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


KERNEL_TAGS = {"fb_iter_finest": 0, "fb_um_finest": 1, "fb_boxv_finest": 2, "fb_boxh_finest": 3,
               "fb_polyexp_finest": 4, "compositor_layer": 5, "post_forward": 6, "hs_sweep": 7,
               "lk_track_finest": 8}


def timer_enable(on: bool):
    check(load().tf_timer_enable(int(bool(on))))


def timer_read(tag: str):
    """-> (total milliseconds, launches) of a tagged kernel since timer_enable(True)."""
    ms, n = C.c_double(), C.c_uint64()
    check(load().tf_timer_read(KERNEL_TAGS[tag], C.byref(ms), C.byref(n)))
    return ms.value, n.value


def launch_count() -> int:
    return int(load().tf_launch_count())
