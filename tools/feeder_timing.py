"""End to end WITH host decode: N MJPG/AVI files decoded by OpenCV's FFmpeg backend.
(a) the reference's shape of work: one file after the other through track(); (b) track_batch: decode threads fill
the page-locked ring while the GPU tracks the previous chunk.  Usage: python tools/feeder_timing.py [N] [frames]"""
import os, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2, pt_import
pkg = pt_import.load()
from tools import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 120
H, W = 480, 640
tmp = tempfile.mkdtemp()
paths, tras = [], []
for s in range(N):
    tra = synth.spiral(0.8 * 240, 3000, (240, 320), seed=s)[:nfr]
    vid = synth.SyntheticVideo(H, W, tra, 25, True, fps=24.0)
    path = os.path.join(tmp, f"v{s}.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 24.0, (W, H), isColor=True)
    for k in range(nfr):
        wr.write(cv2.cvtColor(vid.frame(k), cv2.COLOR_GRAY2BGR))
    wr.release()
    paths.append(path); tras.append(tra)
kw = dict(stop=nfr / 24.0, target_width=25, start_location=pkg.CartesianIndex(240, 320), fps=24)
# decode only (one thread): the floor of the serial path
t0 = time.perf_counter()
for p in paths[:4]:
    c = pkg.CvVideo(p)
    for k in range(nfr): c.frame(k)
dec = (time.perf_counter() - t0) / (4 * nfr)
print(f"decode alone (1 thread, 480x640 MJPG): {dec*1e6:.0f} us/frame")
pkg.track(paths[0], **kw)
t0 = time.perf_counter()
singles = [pkg.track(p, **kw)[1] for p in paths]
t_serial = time.perf_counter() - t0
print(f"serial track() over {N} files: {t_serial*1e3:.0f} ms = {N*nfr/t_serial:.0f} frames/s")
for workers in (4, 8, 16):
    t0 = time.perf_counter()
    ts, ij = pkg.track_batch(paths, decode_workers=workers, **kw)
    t_b = time.perf_counter() - t0
    ok = all(np.array_equal(ij[:, v], singles[v]) for v in range(N))
    print(f"track_batch, {workers:2d} decode threads: {t_b*1e3:.0f} ms = {N*nfr/t_b:.0f} frames/s  identical to serial: {ok}")
