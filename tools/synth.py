"""Synthetic ground-truth videos: the recipe of the reference's own test helper
(test/test-basic-test.jl:19-41, 64-71, 106-113) with a *seeded* jitter and no
codec in the loop: uint8 frames, background 128, one filled disk
of radius target_width÷2 valued 0 (dark) or 255 (light) whose centre follows a
5-loop Archimedean spiral sampled uniformly in arc length.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np


def _arc_len(theta, b):
    """len(θ, b) — test/test-basic-test.jl:19"""
    return b / 2.0 * (theta * np.sqrt(1.0 + theta * theta) + np.arcsinh(theta))


def spiral(r: float, nframes: int, start_ij, seed: int = 0, jitter: float = 1.0) -> np.ndarray:
    """test/test-basic-test.jl:23-33.  Returns (nframes, 2) int64 1-based (row, col)."""
    loops = 5
    a = r / loops / (2.0 * math.pi)
    tmax = loops * 2.0 * math.pi
    total = float(_arc_len(np.float64(tmax), a))
    targets = np.linspace(0.0, total, nframes + 1)[1:]
    # invert the (monotone) arc length by bisection
    lo = np.zeros(nframes)
    hi = np.full(nframes, tmax)
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        big = _arc_len(mid, a) > targets
        hi = np.where(big, mid, hi)
        lo = np.where(big, lo, mid)
    th = 0.5 * (lo + hi)
    noise = np.random.default_rng(seed).standard_normal((nframes, 2)) * jitter
    pts = np.stack([a * th * np.cos(th), a * th * np.sin(th)], axis=1) + noise
    ij = np.rint(pts).astype(np.int64)           # Julia round: ties to even
    return ij - ij[0] + np.asarray(start_ij, np.int64)


def build_trajectory(r: float, fps: float, start_ij, seconds: float = 10.0, seed: int = 0):
    """test/test-basic-test.jl:35-41: ts = 0:1/fps:seconds, one point per frame."""
    nframes = int(math.floor(seconds * fps + 1e-9)) + 1
    ts = np.arange(nframes) / float(fps)
    return ts, spiral(r, nframes, start_ij, seed)


def default_radius(start_ij, H: int, W: int) -> float:
    """test/test-basic-test.jl:110-111: 0.8·min distance from the start to the frame edges."""
    return 0.8 * min(start_ij[0], start_ij[1], H - start_ij[0], W - start_ij[1])


def render_frame(H: int, W: int, centre_ij, disk_radius: int, darker_target: bool = True,
                 out: np.ndarray | None = None, background: int = 128) -> np.ndarray:
    """One frame: test/test-basic-test.jl:65-68 (background Gray(0.5) = 128, disk
    pure black/white).  centre is 1-based (row, col)."""
    if out is None:
        out = np.empty((H, W), np.uint8)
    out[...] = background
    cy, cx = int(centre_ij[0]) - 1, int(centre_ij[1]) - 1
    r = int(disk_radius)
    y0, y1 = max(0, cy - r), min(H, cy + r + 1)
    x0, x1 = max(0, cx - r), min(W, cx + r + 1)
    if y0 < y1 and x0 < x1:
        yy, xx = np.ogrid[y0:y1, x0:x1]
        mask = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        out[y0:y1, x0:x1][mask] = 0 if darker_target else 255
    return out


class SyntheticVideo:
    """An in-memory 'file': frames are rendered on demand from the trajectory.

    `sar` is the sample aspect ratio: stored frames have W = display_W / sar
    columns (test/test-basic-test.jl:72,77 `scale=w÷aspect:h,setsar=aspect`),
    so a displayed column c is stored at column round(c / sar).
    """

    def __init__(self, H: int, W: int, trajectory_ij, target_width: int, darker_target: bool = True,
                 fps: float = 24.0, sar=1, noise_seed: int | None = None, noise_amp: int = 0):
        self.H, self.W = int(H), int(W)
        self.sar = Fraction(sar)
        self.fps = float(fps)
        self.traj = np.asarray(trajectory_ij, np.int64)          # displayed coordinates, 1-based
        self.disk_radius = int(target_width) // 2
        self.darker_target = bool(darker_target)
        self.noise_seed, self.noise_amp = noise_seed, int(noise_amp)

    def __len__(self):
        return len(self.traj)

    @property
    def stored_width(self) -> int:
        return int(self.W // self.sar)

    def stored_centre(self, k: int):
        i, j = self.traj[k]
        return int(i), int(round(Fraction(int(j)) / self.sar))

    def stored_trajectory(self) -> np.ndarray:
        return np.array([self.stored_centre(k) for k in range(len(self))], np.int64)

    def frame(self, k: int, out: np.ndarray | None = None) -> np.ndarray:
        f = render_frame(self.H, self.stored_width, self.stored_centre(k), self.disk_radius,
                         self.darker_target, out)
        if self.noise_amp:
            rng = np.random.default_rng((0 if self.noise_seed is None else self.noise_seed) * 1000003 + k)
            n = rng.integers(-self.noise_amp, self.noise_amp + 1, f.shape)
            np.clip(f.astype(np.int16) + n, 0, 255, out=n)
            f[...] = n.astype(np.uint8)
        return f


def my_partition(n: int, nsegments: int):
    """test/test-basic-test.jl:43-49: nsegments index ranges over 0..n-1 that
    overlap by one frame. Returns a list of (first, last) inclusive, 0-based."""
    edges = np.rint(np.linspace(1, n, nsegments + 1)).astype(int)
    i1 = edges[:-1]
    i2 = list(i1[1:]) + [n]
    return [(int(a) - 1, int(b) - 1) for a, b in zip(i1, i2)]


def make_video(H=100, W=100, target_width=10, darker_target=True, fps=24.0, start_ij=(50, 50), sar=1,
               seconds=10.0, seed=0, **kw) -> SyntheticVideo:
    """The reference test's default case (test/test-basic-test.jl:1-10), any size."""
    r = default_radius(start_ij, H, W)
    _, tra = build_trajectory(r, fps, start_ij, seconds, seed)
    return SyntheticVideo(H, W, tra, target_width, darker_target, fps, sar, **kw)
