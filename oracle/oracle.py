"""ctypes front-end of oracle/libpawsome_oracle.so plus an independent numpy
restatement of the same algorithm (used to cross-check the C code on small
shapes).  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see dog_oracle.c).

All (row, col) indices in this API are 1-based like the reference's
CartesianIndex (src/PawsomeTracker.jl:55-62).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpawsome_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C oracle in place (gcc; a few hundred ms)."""
    src = os.path.join(_HERE, "dog_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libpawsome_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


class _Result(C.Structure):
    _fields_ = [("i", C.c_int), ("j", C.c_int), ("raw_i", C.c_int), ("raw_j", C.c_int),
                ("resp", C.c_double), ("second", C.c_double), ("maxabs", C.c_double)]


@dataclass
class OracleResult:
    i: int          # clamped 1-based row   (what trckr(guess) returns, :61)
    j: int          # clamped 1-based col
    raw_i: int      # unclamped argmax row
    raw_j: int
    resp: float     # maximum response
    second: float   # best response elsewhere in the window
    maxabs: float   # max |R| over the window
    R: np.ndarray | None = None  # (wr, wc) response map if requested

    def near_tie(self, rtol: float = 1e-5) -> bool:
        """top-2 gap below rtol·max|R| ⇒ the argmax is not
        determined at FP32 resolution; such frames are documented, not compared."""
        return (self.resp - self.second) < rtol * max(self.maxabs, 1e-300)


class Oracle:
    def __init__(self, path: str | None = None):
        if path is None:
            path = _SO
            if not os.path.exists(path):
                build()
        L = C.CDLL(path)
        self._L = L
        u8p = C.POINTER(C.c_uint8)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.pto_sigma.restype = C.c_double
        L.pto_sigma.argtypes = [C.c_double]
        L.pto_kernel_len.argtypes = [C.c_double]
        L.pto_default_window.argtypes = [C.c_double]
        L.pto_factors.argtypes = [C.c_double, dp, dp]
        L.pto_dense_kernel.argtypes = [C.c_double, C.c_int, dp]
        L.pto_mode_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t]
        rect = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_double, C.c_int,
                C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_Result), dp]
        L.pto_rect_dense.argtypes = rect
        L.pto_rect_separable.argtypes = rect
        L.pto_tracker_step.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_double, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(_Result), dp]
        L.pto_batch_step_dense.argtypes = [C.POINTER(u8p), C.c_int, C.c_int, C.c_int, C.c_size_t, ip,
                                           C.c_double, C.c_int, C.c_int, C.c_int, ip, ip, dp, C.c_int]
        L.pto_rect_dense_mt.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_double, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_Result)]

    # ---- scalars -----------------------------------------------------------
    def sigma(self, tw: float) -> float:
        return self._L.pto_sigma(tw)

    def kernel_len(self, tw: float) -> int:
        return self._L.pto_kernel_len(tw)

    def default_window(self, tw: float) -> int:
        return self._L.pto_default_window(tw)

    def max_threads(self) -> int:
        return self._L.pto_max_threads()

    def factors(self, tw: float):
        l = self.kernel_len(tw)
        gp = np.empty(l, np.float64)
        gm = np.empty(l, np.float64)
        self._L.pto_factors(tw, gp.ctypes.data_as(C.POINTER(C.c_double)),
                            gm.ctypes.data_as(C.POINTER(C.c_double)))
        return gp, gm

    def dense_kernel(self, tw: float, darker: bool) -> np.ndarray:
        """l×l kernel indexed K[a, b] (a = row offset, b = col offset)."""
        l = self.kernel_len(tw)
        K = np.empty((l, l), np.float64)  # stored [b][a]
        self._L.pto_dense_kernel(tw, int(darker), K.ctypes.data_as(C.POINTER(C.c_double)))
        return K.T.copy()

    # ---- frame helpers -------------------------------------------------------
    @staticmethod
    def _frame(frame: np.ndarray):
        assert frame.dtype == np.uint8 and frame.ndim == 2 and frame.strides[1] == 1
        return frame.ctypes.data_as(C.POINTER(C.c_uint8)), frame.shape[0], frame.shape[1], frame.strides[0]

    def mode(self, frame: np.ndarray) -> int:
        p, H, W, pitch = self._frame(frame)
        return self._L.pto_mode_u8(p, H, W, pitch)

    def _wrap(self, r: _Result, R):
        return OracleResult(r.i, r.j, r.raw_i, r.raw_j, r.resp, r.second, r.maxabs, R)

    def rect(self, frame, fill, tw, darker, y0, x0, wr, wc, dense=True, want_map=False) -> OracleResult:
        """Response over output rows y0..y0+wr-1, cols x0..x0+wc-1 (0-based)."""
        p, H, W, pitch = self._frame(frame)
        r = _Result()
        Rbuf = np.empty((wc, wr), np.float64) if want_map else None
        Rp = Rbuf.ctypes.data_as(C.POINTER(C.c_double)) if want_map else None
        fn = self._L.pto_rect_dense if dense else self._L.pto_rect_separable
        rc = fn(p, H, W, pitch, int(fill), float(tw), int(darker), y0, x0, wr, wc, C.byref(r), Rp)
        if rc != 0:
            raise MemoryError("oracle allocation failed")
        return self._wrap(r, Rbuf.T.copy() if want_map else None)

    def step(self, frame, fill, tw, darker, ws, guess, dense=True, want_map=False) -> OracleResult:
        """(trckr::Tracker)(guess) — src/PawsomeTracker.jl:55-62. ws=(rows, cols)."""
        p, H, W, pitch = self._frame(frame)
        r = _Result()
        wr, wc = 2 * (ws[0] // 2) + 1, 2 * (ws[1] // 2) + 1
        Rbuf = np.empty((wc, wr), np.float64) if want_map else None
        Rp = Rbuf.ctypes.data_as(C.POINTER(C.c_double)) if want_map else None
        rc = self._L.pto_tracker_step(p, H, W, pitch, int(fill), float(tw), int(darker),
                                      int(ws[0]), int(ws[1]), int(guess[0]), int(guess[1]),
                                      int(dense), C.byref(r), Rp)
        if rc != 0:
            raise MemoryError("oracle allocation failed")
        return self._wrap(r, Rbuf.T.copy() if want_map else None)

    def batch_step_dense(self, frames, fills, tw, darker, ws, guess, nthreads=0):
        """One lock-step step over n videos, OpenMP across videos.
        frames: list of HxW u8 arrays (same shape/pitch). guess: (n,2) 1-based.
        Returns (out_ij (n,2) int32, resp (n,), threads_used)."""
        n = len(frames)
        H, W = frames[0].shape
        pitch = frames[0].strides[0]
        u8p = C.POINTER(C.c_uint8)
        ptrs = (u8p * n)(*[f.ctypes.data_as(u8p) for f in frames])
        fills = np.ascontiguousarray(fills, np.int32)
        guess = np.ascontiguousarray(guess, np.int32)
        out = np.empty((n, 2), np.int32)
        resp = np.empty(n, np.float64)
        ip = C.POINTER(C.c_int)
        used = self._L.pto_batch_step_dense(ptrs, n, H, W, pitch, fills.ctypes.data_as(ip), float(tw),
                                            int(darker), int(ws[0]), int(ws[1]),
                                            guess.ctypes.data_as(ip), out.ctypes.data_as(ip),
                                            resp.ctypes.data_as(C.POINTER(C.c_double)), int(nthreads))
        return out, resp, used

    def rect_dense_mt(self, frame, fill, tw, darker, y0, x0, wr, wc, nthreads=0):
        p, H, W, pitch = self._frame(frame)
        r = _Result()
        used = self._L.pto_rect_dense_mt(p, H, W, pitch, int(fill), float(tw), int(darker),
                                         y0, x0, wr, wc, int(nthreads), C.byref(r))
        if used < 0:
            raise MemoryError("oracle allocation failed")
        return self._wrap(r, None), used


# ---------------------------------------------------------------------------
# Independent numpy restatement (no shared code with the C file): used by
# tests/test_oracle.py to cross-check the C oracle on small shapes.
# ---------------------------------------------------------------------------

def numpy_factors(tw: float):
    """KernelFactors.gaussian for σ and σ√2 at l = 4⌈σ√2⌉+1 (Kernel.DoG, call
    site src/PawsomeTracker.jl:43)."""
    sigma = tw / (2.0 * math.sqrt(2.0 * math.log(2.0)))
    sm = sigma * math.sqrt(2.0)
    l = 4 * math.ceil(sm) + 1
    w = l // 2
    x = np.arange(-w, w + 1, dtype=np.float64)
    gp = np.exp(-(x * x) / (2 * sigma * sigma))
    gp /= gp.sum()
    gm = np.exp(-(x * x) / (2 * sm * sm))
    gm /= gm.sum()
    return gp, gm


def numpy_mode_u8(frame: np.ndarray) -> int:
    """StatsBase.mode with its tie rule: the first value to reach the final
    maximum count when scanning column-major (row index fastest)."""
    flat = np.asarray(frame).T.reshape(-1)  # column-major order of the H×W view
    counts = np.bincount(flat, minlength=256)
    M = counts.max()
    tied = np.flatnonzero(counts == M)
    if len(tied) == 1:
        return int(tied[0])
    # each tied value reaches M at its LAST occurrence; earliest such wins
    last = {int(v): int(np.flatnonzero(flat == v)[-1]) for v in tied}
    return min(last, key=last.get)


def numpy_dense_response(frame: np.ndarray, fill: int, tw: float, darker: bool,
                         y0: int, x0: int, wr: int, wc: int) -> np.ndarray:
    """R[yy, xx] = Σ_ab P(y0+yy+a-w, x0+xx+b-w)·K[a,b], P constant-padded with
    fill/255 (PaddedView, :48), K = ±DoG (:42-43).  Float64, numpy summation
    order (not the reference's) — agreement with the C oracle is to ~1e-15."""
    gp, gm = numpy_factors(tw)
    l = len(gp)
    w = l // 2
    K = (np.outer(gp, gp) - np.outer(gm, gm)) * (-1.0 if darker else 1.0)
    H, W = frame.shape
    fr, fc = wr + 2 * w, wc + 2 * w
    P = np.full((fr, fc), fill / 255.0, np.float64)
    ys = np.arange(y0 - w, y0 - w + fr)
    xs = np.arange(x0 - w, x0 - w + fc)
    yi = (ys >= 0) & (ys < H)
    xi = (xs >= 0) & (xs < W)
    if yi.any() and xi.any():
        P[np.ix_(yi, xi)] = frame[np.ix_(ys[yi], xs[xi])].astype(np.float64) / 255.0
    win = np.lib.stride_tricks.sliding_window_view(P, (l, l))  # (wr, wc, l, l)
    return np.einsum("yxab,ab->yx", win, K, optimize=False)
