"""ctypes binding of libpawsome_cuda.so (include/pawsome.h).

The product path has no CPU fallback: if the shared library is missing this
module raises at import time, and every compute call raises PawsomeError when
CUDA is unusable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PAWSOME_CUDA_LIB: load another build of the same ABI (tools/phase_timing.py loads the -DPT_PROBES build)
LIB_PATH = os.environ.get("PAWSOME_CUDA_LIB") or os.path.join(_HERE, "libpawsome_cuda.so")

PT_OK = 0
PT_PIX_U8 = 0
PT_PIX_F32 = 1
STATUS = {0: "PT_OK", -1: "PT_ERR_ARG", -2: "PT_ERR_CUDA", -3: "PT_ERR_NOMEM",
          -4: "PT_ERR_STATE", -5: "PT_ERR_UNSUPPORTED"}


class PawsomeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the CUDA extension is not built and there is no CPU fallback. "
        "Run `python -c 'import __graft_entry__ as g; g.build()'` (or `make -C pawsometracker.jl_b200/csrc`).")

lib = C.CDLL(LIB_PATH)

_vp = C.c_void_p
_i32p = C.POINTER(C.c_int32)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)

# every symbol include/pawsome.h declares, with its signature
SIGNATURES = {
    "pt_version": (C.c_int, []),
    "pt_last_error": (C.c_char_p, []),
    "pt_device_count": (C.c_int, []),
    "pt_preferred_batch": (C.c_int, [C.c_int]),
    "pt_sigma": (C.c_double, [C.c_double]),
    "pt_kernel_len": (C.c_int, [C.c_double]),
    "pt_default_window": (C.c_int, [C.c_double]),
    "pt_factors_f32": (C.c_int, [C.c_double, C.c_int, _fp, _fp, _fp, _fp]),
    "pt_batch_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.POINTER(_vp)]),
    "pt_batch_destroy": (None, [_vp]),
    "pt_batch_set_window": (C.c_int, [_vp, C.c_int, C.c_int]),
    "pt_batch_set_frames": (C.c_int, [_vp, C.POINTER(_vp), C.c_size_t]),
    "pt_batch_bind_device_frames": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t]),
    "pt_batch_compute_fill": (C.c_int, [_vp, _ip]),
    "pt_batch_set_fill": (C.c_int, [_vp, _ip]),
    "pt_batch_set_guess": (C.c_int, [_vp, _i32p]),
    "pt_batch_step": (C.c_int, [_vp, _i32p, _i32p, _i32p, _fp]),
    "pt_batch_track_device": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, _i32p, _fp]),
    "pt_batch_track_host": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_size_t, C.c_int, _i32p, _fp]),
    "pt_batch_response_map": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _fp]),
    "pt_batch_track_device_async": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, _vp]),
    "pt_batch_read_track": (C.c_int, [_vp, C.c_int, _i32p, _fp]),
    "pt_batch_launch_count": (C.c_longlong, [_vp]),
    "pt_batch_kernel_name": (C.c_char_p, [_vp]),
    "pt_batch_last_kernel": (C.c_char_p, [_vp]),
    "pt_batch_downscale": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_uint8)]),
    "pt_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "pt_host_free": (C.c_int, [_vp]),
    "pt_batch_stream": (_vp, [_vp]),
    "pt_tracker_create": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(_vp)]),
    "pt_tracker_destroy": (None, [_vp]),
    "pt_tracker_set_frame": (C.c_int, [_vp, _vp, C.c_size_t]),
    "pt_tracker_compute_fill": (C.c_int, [_vp, _ip]),
    "pt_tracker_set_fill": (C.c_int, [_vp, C.c_int]),
    "pt_tracker_step": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _ip, _fp]),
    "pt_tracker_step_host": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, C.c_int, _ip, _ip, _fp]),
    "pt_tracker_batch": (_vp, [_vp]),
    "pt_batch_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "pt_batch_rect_argmax": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _ip, _fp]),
    "pt_batch_rect_argmax_all": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _fp, C.c_int]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here = the .so does not match the header
    _f.restype = _res
    _f.argtypes = _args


def last_error() -> str:
    return lib.pt_last_error().decode("utf-8", "replace")


def check(code: int) -> int:
    if code < 0:
        raise PawsomeError(code, last_error())
    return code
