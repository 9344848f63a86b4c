"""Host-side mirror of the reference's `Tracker` (src/PawsomeTracker.jl:32-62)
and its batched counterpart, over the C ABI of libpawsome_cuda.so.

Indices are 1-based (row, col) exactly as in the reference so that the parity
tests read like the reference's own code.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PT_PIX_F32, PT_PIX_U8, check, lib

_i32p = C.POINTER(C.c_int32)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def get_sigma(target_width: float) -> float:
    """src/PawsomeTracker.jl:30"""
    return lib.pt_sigma(float(target_width))


def kernel_len(target_width: float) -> int:
    """Length l of Kernel.DoG(σ) (call site src/PawsomeTracker.jl:43)."""
    return check(lib.pt_kernel_len(float(target_width)))


def guess_window_size(target_width: float) -> int:
    """src/PawsomeTracker.jl:64-68"""
    return lib.pt_default_window(float(target_width))


def fix_window_size(ws):
    """src/PawsomeTracker.jl:70-72: (w, h) → (h, w); l → (l, l)."""
    if isinstance(ws, (tuple, list)):
        if len(ws) != 2:
            raise ValueError("window_size must be an Int or a (w, h) tuple")
        return (int(ws[1]), int(ws[0]))
    return (int(ws), int(ws))


def factors_f32(target_width: float, darker_target: bool):
    l = kernel_len(target_width)
    arrs = [np.empty(l, np.float32) for _ in range(4)]
    check(lib.pt_factors_f32(float(target_width), int(darker_target), *[a.ctypes.data_as(_fp) for a in arrs]))
    return tuple(arrs)


def _pixel_of(dtype) -> int:
    if dtype == np.uint8:
        return PT_PIX_U8
    if dtype == np.float32:
        return PT_PIX_F32
    raise TypeError(f"frames must be uint8 (Gray{{N0f8}}) or float32, got {dtype}")


def _check_frame(img: np.ndarray, H: int, W: int, pixel: int):
    if img.ndim != 2 or img.shape != (H, W):
        raise ValueError(f"DimensionMismatch: frame is {img.shape}, tracker was built for {(H, W)}")
    if _pixel_of(img.dtype) != pixel:
        raise TypeError("frame dtype differs from the tracker's pixel type")
    if img.strides[1] != img.itemsize:
        raise ValueError("frame rows must be contiguous (row-major)")
    return img.strides[0] // img.itemsize


class TrackerBatch:
    """n independent `Tracker`s of identical geometry advanced in lock-step,
    one CTA group per (video, window) — the batched entry the reference lacks."""

    def __init__(self, n, frame_size, target_width, window_size, darker_target,
                 dtype=np.uint8, device: int = 0):
        H, W = int(frame_size[0]), int(frame_size[1])
        self.n, self.sz = int(n), (H, W)
        self.target_width = float(target_width)
        self.window_size = (int(window_size[0]), int(window_size[1]))
        self.radii = (self.window_size[0] // 2, self.window_size[1] // 2)
        self.darker_target = bool(darker_target)
        self.pixel = _pixel_of(np.dtype(dtype))
        self.dtype = np.dtype(dtype)
        self.device = int(device)
        h = C.c_void_p()
        check(lib.pt_batch_create(self.n, H, W, self.target_width, self.window_size[0], self.window_size[1],
                                  int(self.darker_target), self.pixel, self.device, C.byref(h)))
        self._h = h

    # -- lifetime ------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib.pt_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- configuration ---------------------------------------------------------
    def set_window(self, window_size):
        check(lib.pt_batch_set_window(self._h, int(window_size[0]), int(window_size[1])))
        self.window_size = (int(window_size[0]), int(window_size[1]))
        self.radii = (self.window_size[0] // 2, self.window_size[1] // 2)

    def set_option(self, name: str, value: int):
        """Per-handle tuning / debugging knob (pt_batch_set_option): "window45", "rect45", "rot", "skew",
        "r45_chunks", "generic_target", "mode_slow", "zero_copy", "host_lanes", "cluster", "bulk", "wide"."""
        check(lib.pt_batch_set_option(self._h, name.encode(), int(value)))

    def set_frames(self, frames):
        """frames: sequence of n HxW arrays (host)."""
        if len(frames) != self.n:
            raise ValueError(f"expected {self.n} frames, got {len(frames)}")
        pitches = {_check_frame(f, self.sz[0], self.sz[1], self.pixel) for f in frames}
        if len(pitches) != 1:
            raise ValueError("all frames of a step must share one pitch")
        ptrs = (C.c_void_p * self.n)(*[f.ctypes.data for f in frames])
        check(lib.pt_batch_set_frames(self._h, ptrs, pitches.pop()))

    def bind_device_frames(self, dev_ptr: int, frame_stride: int, pitch: int):
        check(lib.pt_batch_bind_device_frames(self._h, C.c_void_p(dev_ptr), frame_stride, pitch))

    def compute_fill(self) -> np.ndarray:
        out = np.empty(self.n, np.int32)
        check(lib.pt_batch_compute_fill(self._h, out.ctypes.data_as(_ip)))
        return out

    def set_fill(self, fills):
        f = np.ascontiguousarray(np.broadcast_to(np.asarray(fills, np.int32), (self.n,)))
        check(lib.pt_batch_set_fill(self._h, f.ctypes.data_as(_ip)))

    def set_guess(self, guess):
        g = np.ascontiguousarray(guess, np.int32).reshape(self.n, 2)
        check(lib.pt_batch_set_guess(self._h, g.ctypes.data_as(_i32p)))

    # -- compute ---------------------------------------------------------------
    def step(self, guess=None, want_raw: bool = False):
        """One `trckr(guess)` per video (src/PawsomeTracker.jl:55-62).
        Returns (ij (n,2) int32 1-based clamped, resp (n,) float32[, raw (n,2)])."""
        g = None
        if guess is not None:
            g = np.ascontiguousarray(guess, np.int32).reshape(self.n, 2)
        out = np.empty((self.n, 2), np.int32)
        raw = np.empty((self.n, 2), np.int32)
        resp = np.empty(self.n, np.float32)
        check(lib.pt_batch_step(self._h, g.ctypes.data_as(_i32p) if g is not None else None,
                                out.ctypes.data_as(_i32p), raw.ctypes.data_as(_i32p), resp.ctypes.data_as(_fp)))
        return (out, resp, raw) if want_raw else (out, resp)

    def track_device(self, dev_ptr: int, step_stride: int, frame_stride: int, pitch: int, T: int):
        out = np.empty((T, self.n, 2), np.int32)
        resp = np.empty((T, self.n), np.float32)
        check(lib.pt_batch_track_device(self._h, C.c_void_p(dev_ptr), step_stride, frame_stride, pitch, T,
                                        out.ctypes.data_as(_i32p), resp.ctypes.data_as(_fp)))
        return out, resp

    def track_device_async(self, dev_ptr: int, step_stride: int, frame_stride: int, pitch: int, T: int,
                           stream: int = 0):
        check(lib.pt_batch_track_device_async(self._h, C.c_void_p(dev_ptr), step_stride, frame_stride, pitch, T,
                                              C.c_void_p(stream) if stream else None))

    def read_track(self, T: int):
        out = np.empty((T, self.n, 2), np.int32)
        resp = np.empty((T, self.n), np.float32)
        check(lib.pt_batch_read_track(self._h, T, out.ctypes.data_as(_i32p), resp.ctypes.data_as(_fp)))
        return out, resp

    def track_host(self, frames, mode: str = "footprint"):
        """frames[t][v]: host arrays. Returns (ij (T,n,2), resp (T,n))."""
        T = len(frames)
        flat = []
        pitch = None
        for step in frames:
            if len(step) != self.n:
                raise ValueError(f"expected {self.n} frames per step")
            for f in step:
                p = _check_frame(f, self.sz[0], self.sz[1], self.pixel)
                if pitch is None:
                    pitch = p
                elif p != pitch:
                    raise ValueError("all frames must share one pitch")
                flat.append(f.ctypes.data)
        return self.track_host_ptrs(flat, T, pitch, mode)

    def make_ptr_table(self, ptrs):
        """Pack frame addresses (T*n, step-major) into the C pointer table track_host_ptrs takes."""
        return (C.c_void_p * len(ptrs))(*ptrs)

    def track_host_ptrs(self, ptrs, T: int, pitch: int, mode: str = "footprint"):
        m = {"footprint": 0, "frames": 1}[mode]
        arr = ptrs if isinstance(ptrs, C.Array) else (C.c_void_p * (T * self.n))(*ptrs)
        out = np.empty((T, self.n, 2), np.int32)
        resp = np.empty((T, self.n), np.float32)
        check(lib.pt_batch_track_host(self._h, arr, T, pitch, m, out.ctypes.data_as(_i32p), resp.ctypes.data_as(_fp)))
        return out, resp

    def response_map(self, v: int, guess) -> np.ndarray:
        wr, wc = 2 * self.radii[0] + 1, 2 * self.radii[1] + 1
        out = np.empty((wr, wc), np.float32)
        check(lib.pt_batch_response_map(self._h, int(v), int(guess[0]), int(guess[1]), out.ctypes.data_as(_fp)))
        return out

    def rect_argmax(self, v: int, y0: int, x0: int, wr: int, wc: int):
        oi, oj, ri, rj = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        r = C.c_float()
        check(lib.pt_batch_rect_argmax(self._h, int(v), int(y0), int(x0), int(wr), int(wc), C.byref(oi), C.byref(oj),
                                       C.byref(ri), C.byref(rj), C.byref(r)))
        return (oi.value, oj.value), (ri.value, rj.value), r.value

    def downscale(self, out_h: int, out_w: int) -> np.ndarray:
        """`imresize!` of the current frame of every video to (out_h, out_w) on the device (src/diagnose.jl:33)."""
        out = np.empty((self.n, int(out_h), int(out_w)), np.uint8)
        check(lib.pt_batch_downscale(self._h, int(out_h), int(out_w), out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def rect_argmax_all(self, y0: int, x0: int, wr: int, wc: int, readback: bool = True):
        """The same output rectangle on the current frame of every video, one launch.  Returns
        (ij [n,2], raw_ij [n,2], resp [n]); with readback=False only enqueues the launch."""
        if not readback:
            check(lib.pt_batch_rect_argmax_all(self._h, int(y0), int(x0), int(wr), int(wc), None, None, 1))
            return None
        out = np.empty((self.n, 4), np.int32)
        resp = np.empty(self.n, np.float32)
        check(lib.pt_batch_rect_argmax_all(self._h, int(y0), int(x0), int(wr), int(wc),
                                           out.ctypes.data_as(C.POINTER(C.c_int32)), resp.ctypes.data_as(_fp), 0))
        return out[:, :2].copy(), out[:, 2:].copy(), resp

    # -- introspection -----------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return lib.pt_batch_launch_count(self._h)

    @property
    def kernel_name(self) -> str:
        return lib.pt_batch_kernel_name(self._h).decode()

    @property
    def last_kernel(self) -> str:
        """Kernel the most recent step / track launch ran."""
        return lib.pt_batch_last_kernel(self._h).decode()

    @property
    def stream(self) -> int:
        return lib.pt_batch_stream(self._h) or 0


class Tracker:
    """`Tracker(img, target_width, window_size, darker_target)` —
    src/PawsomeTracker.jl:39-52.  `trckr.img` is the writable host frame the
    decoder fills (`trckr.img.data` in the reference, :166); calling the
    tracker with a guess runs :55-62 on the GPU, copying only the window's
    footprint of that frame."""

    def __init__(self, img: np.ndarray, target_width, window_size, darker_target, device: int = 0, _fillvalue=None):
        # (_fillvalue: the mode of THIS img when the caller has just computed it — the auto-detect start builds two
        # trackers on the same first frame, :103-105; saves a whole-frame upload and a mode pass, same value)
        if img.ndim != 2:
            raise ValueError("img must be a 2-D grayscale frame")
        self._call = None
        self.img = np.ascontiguousarray(img)
        self.sz = tuple(int(s) for s in self.img.shape)
        ws = (int(window_size[0]), int(window_size[1]))
        self.radii = (ws[0] // 2, ws[1] // 2)
        self.target_width = float(target_width)
        self.darker_target = bool(darker_target)
        self._batch = TrackerBatch(1, self.sz, target_width, ws, darker_target, dtype=self.img.dtype, device=device)
        # fillvalue = mode(_img) of THIS frame, reused for all later frames (:47)
        if _fillvalue is None:
            self._batch.set_frames([self.img])
            self.fillvalue = int(self._batch.compute_fill()[0])
        else:
            self.fillvalue = int(_fillvalue)
            self._batch.set_fill([self.fillvalue])
        self.last_response = float("nan")

    @property
    def img(self) -> np.ndarray:
        """The writable host frame (`trckr.img.data`, :166).  Fill it in place (`read(out=trckr.img)`); assigning a new
        array is allowed too and re-targets the per-call argument buffers."""
        return self._img

    @img.setter
    def img(self, arr):
        self._img = arr
        c = self._call
        if c is not None:                   # re-target the cached argument buffers (the ctypes arrays are kept)
            c[1][0] = arr.ctypes.data
            c[2] = _check_frame(arr, self.sz[0], self.sz[1], self._batch.pixel)

    def __call__(self, guess):
        # a tracker is a batch of one: footprint streaming of the current host frame.  This is the per-frame
        # call of the reference's loop (:167), so the ctypes argument buffers are built once and reused.
        c = self._call
        if c is None:
            pitch = _check_frame(self.img, self.sz[0], self.sz[1], self._batch.pixel)
            c = self._call = [(C.c_int32 * 2)(), (C.c_void_p * 1)(self.img.ctypes.data), pitch,
                              (C.c_int32 * 2)(), (C.c_float * 1)()]
        g, ptrs, pitch, out, resp = c
        g[0], g[1] = int(guess[0]), int(guess[1])
        h = self._batch._h
        check(lib.pt_batch_set_guess(h, g))
        check(lib.pt_batch_track_host(h, ptrs, 1, pitch, 0, out, resp))
        self.last_response = float(resp[0])
        return (out[0], out[1])

    def track_frames(self, frames, guess):
        """The frame loop `indices[k] = trckr(indices[k-1])` (src/PawsomeTracker.jl:163-169) over a list of host frames
        in ONE library call (pt_batch_track_host): frames in page-locked memory (pt_host_alloc / PinnedArray — where a
        decoder should write them) are read in place by one chained launch; pageable frames take the per-step
        footprint path inside the library.  Returns (ij (T, 2) int32, resp (T,) float32)."""
        T = len(frames)
        pitches = {_check_frame(f, self.sz[0], self.sz[1], self._batch.pixel) for f in frames}
        if len(pitches) != 1:
            raise ValueError("all frames of a chunk must share one pitch")
        ptrs = (C.c_void_p * T)(*[f.__array_interface__["data"][0] for f in frames])
        g = (C.c_int32 * 2)(int(guess[0]), int(guess[1]))
        out = np.empty((T, 2), np.int32)
        resp = np.empty(T, np.float32)
        h = self._batch._h
        check(lib.pt_batch_set_guess(h, g))
        check(lib.pt_batch_track_host(h, ptrs, T, pitches.pop(), 0, out.ctypes.data_as(_i32p), resp.ctypes.data_as(_fp)))
        self.last_response = float(resp[-1])
        return out, resp

    def track_addresses(self, addrs: np.ndarray, pitch: int, guess):
        """track_frames for frames given by ADDRESS (an int64 array, one host frame each, rows `pitch` elements apart):
        the caller vouches for shape and dtype.  One library call for the whole block."""
        addrs = np.ascontiguousarray(addrs, np.int64)
        T = int(addrs.size)
        g = (C.c_int32 * 2)(int(guess[0]), int(guess[1]))
        out = np.empty((T, 2), np.int32)
        resp = np.empty(T, np.float32)
        h = self._batch._h
        check(lib.pt_batch_set_guess(h, g))
        check(lib.pt_batch_track_host(h, addrs.ctypes.data_as(C.POINTER(C.c_void_p)), T, int(pitch), 0,
                                      out.ctypes.data_as(_i32p), resp.ctypes.data_as(_fp)))
        self.last_response = float(resp[-1])
        return out, resp

    def step_resident(self, guess):
        """Same result with the whole frame uploaded to HBM first (the
        `read!` + step of the reference's loop, :166-167)."""
        self._batch.set_frames([self.img])
        out, resp = self._batch.step([[int(guess[0]), int(guess[1])]])
        self.last_response = float(resp[0])
        return (int(out[0, 0]), int(out[0, 1]))

    def response_map(self, guess) -> np.ndarray:
        self._batch.set_frames([self.img])
        return self._batch.response_map(0, guess)

    def downscaled(self, out_h: int, out_w: int) -> np.ndarray:
        """`imresize!` of the current host frame (uploaded whole) on the device — diagnostics only."""
        self._batch.set_frames([self.img])
        return self._batch.downscale(out_h, out_w)[0]

    def set_option(self, name: str, value: int):
        self._batch.set_option(name, value)

    def close(self):
        self._batch.close()
