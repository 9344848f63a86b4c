"""ctypes binding of libpawsome_bench.so (include/pawsome_bench.h): measurement helpers for bench.py and tools/ —
the FP32 peak of the device (roofline denominator), an issue-port probe and an L2 flush.  Deliberately outside the
product package: libpawsome_cuda.so exports the DoG-window + argmax path and nothing else."""
import ctypes as C
import os

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH_LIB_PATH = os.path.join(_ROOT, "pawsometracker.jl_b200", "libpawsome_bench.so")

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(BENCH_LIB_PATH):
            raise ImportError(f"{BENCH_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(BENCH_LIB_PATH)
        L.ptb_measure_fp32_peak.restype = C.c_int
        L.ptb_measure_fp32_peak.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.ptb_probe_ffma2_issue.restype = C.c_int
        L.ptb_probe_ffma2_issue.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.ptb_flush_l2.restype = C.c_int
        L.ptb_flush_l2.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.ptb_time_chain.restype = C.c_int
        L.ptb_time_chain.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                     C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def time_chain(product_lib, batch, segments, step_stride: int, frame_stride: int, pitch: int, stream: int) -> float:
    """Device time (ms) of pt_batch_track_device_async over `segments` = [(device pointer, steps), ...] on `stream`:
    event, launches, event issued back to back from C (ptb_time_chain); returns after the second event completed."""
    fn = C.cast(product_lib.pt_batch_track_device_async, C.c_void_p)
    nseg = len(segments)
    bases = (C.c_void_p * nseg)(*[int(p) for p, _ in segments])
    Ts = (C.c_int * nseg)(*[int(t) for _, t in segments])
    ms = C.c_double()
    rc = load().ptb_time_chain(fn, batch._h, nseg, bases, Ts, step_stride, frame_stride, pitch,
                               C.c_void_p(stream) if stream else None, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"ptb_time_chain failed ({rc})")
    return ms.value


def fp32_peak(device: int, packed: int, reps: int = 5) -> float:
    """TFLOP/s: packed 0 scalar FFMA, 1 fma.rn.f32x2 (2 flops per lane-op), 2 add.rn.f32x2 (1 flop per lane-op)."""
    v = C.c_double()
    rc = load().ptb_measure_fp32_peak(int(device), int(packed), int(reps), C.byref(v))
    if rc != 0:
        raise RuntimeError(f"ptb_measure_fp32_peak failed ({rc})")
    return v.value


def flush_l2(ptr: int, nbytes: int, stream: int):
    rc = load().ptb_flush_l2(C.c_void_p(ptr), int(nbytes), C.c_void_p(stream) if stream else None)
    if rc != 0:
        raise RuntimeError(f"ptb_flush_l2 failed ({rc})")
