"""Wall-clock timings of the BASELINE configs through the public API (track / Tracker), one GPU.
Frames are pre-rendered into host memory (decode is out of scope), so the numbers are the tracker's own
per-frame cost as a user of the drop-in API sees it.  Usage: python tools/config_timings.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pt_import
pkg = pt_import.load()
from tools import synth


def timed(label, fn, reps=3):
    best = 1e9
    out = None
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    return best, out


def config_track(label, H, W, nfr, tw, darker, start_location, window_size=None):
    start = (H // 2, W // 2)
    tra = synth.spiral(0.8 * min(H, W) / 2, 3000, start, seed=0)[:nfr]          # the 3000-frame spiral of the configs
    vid = synth.SyntheticVideo(H, W, tra, tw, darker, fps=24.0)
    frames = np.stack([vid.frame(k) for k in range(nfr)])
    ref = None
    for where in ("pageable", "page-locked"):
        # page-locked: the decoder's output buffer is pt_host_alloc memory (PinnedArray) — the kernels read the
        # footprints in place (zero-copy), one chained launch per 64-frame chunk; pageable: the library gathers each
        # footprint into pinned staging per step
        pin = None
        if where == "page-locked":
            pin = pkg.PinnedArray(frames.shape, np.uint8)
            pin.array[...] = frames
        av = pkg.ArrayVideo(pin.array if pin else frames, fps=24.0)
        dt, (ts, ij) = timed(label, lambda: pkg.track(av, stop=nfr / 24.0, target_width=tw, start_location=start_location,
                                                       window_size=window_size, darker_target=darker, fps=24))
        err = np.sqrt(np.mean(np.sum((ij - tra[:len(ij)]) ** 2, axis=1)))
        same = "" if ref is None else f", identical to pageable: {bool(np.array_equal(ij, ref))}"
        ref = ij if ref is None else ref
        print(f"{label} [{where} frames]: {len(ij)} frames in {dt*1e3:.1f} ms = {len(ij)/dt:.0f} frames/s "
              f"({dt/len(ij)*1e6:.1f} us/frame), RMSE {err:.2f} px{same}", flush=True)
        if pin:
            del av
            pin.close()
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
config_track("config 2: 1080p, 1000 of 3000 frames, tw=25, start missing (auto-detect)", 1080, 1920, 1000, 25, True, None)
config_track("config 4: 4K, 60 frames, light target tw=100, window 401", 2160, 3840, 60, 100, False, None, 401)
config_track("config 4b: 4K, 60 frames, light target tw=100, default window", 2160, 3840, 60, 100, False, None)

# config 5: segmented multi-file video, SAR = 2, non-zero start, fps resampling (serial chain and parallel chains)
H5, Wd5, sar5, src_fps, fps5 = 1080, 1920, 2, 24.0, 12.0
_, tra5 = synth.build_trajectory(0.8 * 540, src_fps, (540, 960), seconds=12.0, seed=0)
parts = synth.my_partition(len(tra5), 3)
segs = [pkg.ArrayVideo(np.stack([synth.SyntheticVideo(H5, Wd5, tra5[a:b + 1], 25, True, fps=src_fps, sar=sar5).frame(k)
                                 for k in range(b - a + 1)]), fps=src_fps, sar=sar5) for a, b in parts]
seg_start = [0.25, 0.0, 0.0]
seg_stop = [(b - a + 1) / src_fps for a, b in parts]
x0, y0 = int(tra5[6, 1]), int(tra5[6, 0])
kw5 = dict(start=seg_start, stop=seg_stop, target_width=25, darker_target=True, fps=fps5)
dt, (ts5, ij5) = timed("c5", lambda: pkg.track(segs, start_location=[(x0, y0), None, None], **kw5))
print(f"config 5: 3 segments 1080x960 (SAR 2), start 0.25 s, fps 24->12, chained: {len(ij5)} frames in {dt*1e3:.1f} ms = "
      f"{len(ij5)/dt:.0f} frames/s")
locs = [(x0, y0)] + [(int(tra5[a, 1]), int(tra5[a, 0])) for a, _ in parts[1:]]
dt, (ts5p, ij5p) = timed("c5p", lambda: pkg.track(segs, start_location=locs, parallel=True, **kw5))
print(f"config 5, every segment with its own start_location, parallel chains: {len(ij5p)} frames in {dt*1e3:.1f} ms = "
      f"{len(ij5p)/dt:.0f} frames/s")
