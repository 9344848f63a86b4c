"""Generates tests/golden/oracle_cases.npz from the CPU oracle.

PARITY UNPINNED: these are regression pins of the restatement (and the
vectors the GPU parity tests replay), NOT outputs of the Julia reference —
julia / ffmpeg / ImageFiltering.jl are not available in this image
(SURVEY §8c).  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import Oracle, build  # noqa: E402


def disk(H, W, cy, cx, r, val, bg=128):
    f = np.full((H, W), bg, np.uint8)
    yy, xx = np.ogrid[0:H, 0:W]
    f[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = val
    return f


def main():
    build()
    o = Oracle()
    rng = np.random.default_rng(20261018)
    cases = []
    # (frame, tw, darker, ws_rows, ws_cols, guess_i, guess_j)
    cases.append((disk(120, 160, 59, 79, 12, 0), 25, 1, 45, 45, 55, 86))            # config-1-like, dark
    cases.append((disk(120, 160, 30, 140, 12, 255), 25, 0, 45, 45, 36, 136))        # light target
    cases.append((disk(96, 128, 4, 5, 5, 0), 10, 1, 21, 21, 8, 9))                  # corner: window leaves the frame
    cases.append((disk(96, 128, 90, 120, 5, 0), 10, 1, 21, 31, 92, 118))            # non-square window, bottom-right
    noisy = disk(100, 100, 49, 49, 5, 0)
    noisy = np.clip(noisy.astype(int) + rng.integers(-20, 21, noisy.shape), 0, 255).astype(np.uint8)
    cases.append((noisy, 10, 1, 21, 21, 47, 53))                                     # noise (H.264-like dirt)
    cases.append((rng.integers(0, 256, (64, 72)).astype(np.uint8), 7, 1, 15, 17, 30, 40))   # pure noise
    cases.append((disk(200, 200, 99, 99, 16, 0), 33, 1, 61, 61, 95, 104))           # another kernel length (l=85)
    out = {"n": np.int64(len(cases))}
    for c, (f, tw, darker, wsr, wsc, gi, gj) in enumerate(cases):
        fill = o.mode(f)
        r = o.step(f, fill, tw, bool(darker), (wsr, wsc), (gi, gj), dense=True, want_map=True)
        out[f"frame{c}"] = f
        out[f"par{c}"] = np.array([tw, darker, wsr, wsc, gi, gj, fill], np.float64)
        out[f"res{c}"] = np.array([r.i, r.j, r.raw_i, r.raw_j, r.resp, r.second, r.maxabs], np.float64)
        out[f"map{c}"] = r.R
        print(c, f.shape, "tw", tw, "→", (r.i, r.j), "resp %.6g" % r.resp, "near_tie", r.near_tie())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_cases.npz"), **out)


if __name__ == "__main__":
    main()
