"""BASELINE config 4 through the public track() at two lengths: per-frame cost in the steady state (slope) vs the start-up
(two trackers, the 4K auto-detect pass).  Page-locked frames.  Usage: python tools/config4_slope.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pt_import
pkg = pt_import.load()
from tools import synth
H, W, tw = 2160, 3840, 100
for nfr in (40, 160):
    tra = synth.spiral(0.8 * min(H, W) / 2, 3000, (H // 2, W // 2), seed=0)[:nfr]
    vid = synth.SyntheticVideo(H, W, tra, tw, False, fps=24.0)
    pin = pkg.PinnedArray((nfr, H, W), np.uint8)
    for k in range(nfr): pin.array[k] = vid.frame(k)
    av = pkg.ArrayVideo(pin.array, fps=24.0)
    for ws in (401, None):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            ts, ij = pkg.track(av, stop=nfr / 24.0, target_width=tw, start_location=None, window_size=ws, darker_target=False, fps=24)
            best = min(best, time.perf_counter() - t0)
        print(f"nfr={nfr} ws={ws}: {best*1e3:.2f} ms total, {best/nfr*1e6:.1f} us/frame", flush=True)
    del av; pin.close()
