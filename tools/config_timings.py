"""Wall-clock timings of the BASELINE configs through the public API (track / Tracker), one GPU.
Frames are pre-rendered into host memory (decode is out of scope), so the numbers are the tracker's own
per-frame cost as a user of the drop-in API sees it.  Usage: python tools/config_timings.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pt_import
pkg = pt_import.load()


def timed(label, fn, reps=3):
    best = 1e9
    out = None
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    return best, out


def config_track(label, H, W, nfr, tw, darker, start_location, window_size=None):
    start = (H // 2, W // 2)
    tra = pkg.spiral(0.8 * min(H, W) / 2, 3000, start, seed=0)[:nfr]          # the 3000-frame spiral of the configs
    vid = pkg.SyntheticVideo(H, W, tra, tw, darker, fps=24.0)
    frames = np.stack([vid.frame(k) for k in range(nfr)])
    av = pkg.ArrayVideo(frames, fps=24.0)
    dt, (ts, ij) = timed(label, lambda: pkg.track(av, stop=nfr / 24.0, target_width=tw, start_location=start_location,
                                                   window_size=window_size, darker_target=darker, fps=24))
    err = np.sqrt(np.mean(np.sum((ij - tra[:len(ij)]) ** 2, axis=1)))
    print(f"{label}: {len(ij)} frames in {dt*1e3:.1f} ms = {len(ij)/dt:.0f} frames/s ({dt/len(ij)*1e6:.1f} us/frame), RMSE {err:.2f} px")
    return frames, tra


# per-call latency of trckr(guess) on a host frame (pageable numpy memory)
f = np.full((1080, 1920), 128, np.uint8)
yy, xx = np.ogrid[0:1080, 0:1920]
f[(yy - 500) ** 2 + (xx - 900) ** 2 <= 144] = 0
trk = pkg.Tracker(f, 25, (45, 45), True)
for _ in range(20): trk((498, 903))
t0 = time.perf_counter()
for _ in range(500): r = trk((498, 903))
dt = (time.perf_counter() - t0) / 500
print(f"Tracker.__call__ (host frame, footprint path), 1080p tw=25: {dt*1e6:.1f} us/call result {r}")
t0 = time.perf_counter()
for _ in range(50): r = trk.step_resident((498, 903))
dt = (time.perf_counter() - t0) / 50
print(f"Tracker.step_resident (whole-frame upload + step): {dt*1e6:.1f} us/call result {r}")
trk.close()

config_track("config 1: 480x640, 300 frames, tw=25, start given", 480, 640, 300, 25, True, pkg.CartesianIndex(240, 320))
config_track("config 2: 1080p, 600 of 3000 frames, tw=25, start missing (auto-detect)", 1080, 1920, 600, 25, True, None)
config_track("config 4: 4K, 60 frames, light target tw=100, window 401", 2160, 3840, 60, 100, False, None, 401)
config_track("config 4b: 4K, 60 frames, light target tw=100, default window", 2160, 3840, 60, 100, False, None)
