"""One small invocation of every kernel, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pt_import
pkg = pt_import.load()
rng = np.random.default_rng(0)
H, W = 96, 128
# window45 (static, T=1 and chained T=3 via device frames), rot (n=240, T=3), generic (tw=10), march (big window), mode, downscale
import ctypes as C
def dev_frames(b, frames):
    b.set_frames(frames)
for n, T in ((3, 1), (240, 3)):
    frames = rng.integers(0, 256, (T, n, H, W)).astype(np.uint8)
    with pkg.TrackerBatch(n, (H, W), 25, (45, 45), True) as b:
        b.set_frames(list(frames[0])); b.compute_fill()
        g = np.stack([rng.integers(1, H, n), rng.integers(1, W, n)], axis=-1)
        ij, r = b.step(g)
        print("window45 step", n, ij[0], b.last_kernel)
        if T > 1:
            import torch
            dev = torch.from_numpy(frames).cuda(); torch.cuda.synchronize()
            b.bind_device_frames(dev.data_ptr(), H * W, W)
            b.set_guess(g)
            ij2, _ = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
            print("chained", n, T, ij2[-1, 0], b.last_kernel)
with pkg.TrackerBatch(2, (H, W), 10, (21, 33), False) as b:
    fr = list(rng.integers(0, 256, (2, H, W)).astype(np.uint8))
    b.set_frames(fr); b.compute_fill()
    print("generic", b.step([[40, 50], [3, 120]])[0], b.last_kernel)
with pkg.TrackerBatch(2, (200, 260), 25, (131, 141), True) as b:
    fr = list(rng.integers(0, 256, (2, 200, 260)).astype(np.uint8))
    b.set_frames(fr); b.compute_fill()
    print("march", b.step([[100, 130], [20, 250]])[0], b.last_kernel)
    print("downscale", b.downscale(50, 65).shape)
f = np.zeros((64, 64), np.uint8); f[:, 32:] = 200
with pkg.TrackerBatch(1, f.shape, 10, (21, 21), True) as b:
    b.set_frames([f]); print("mode tie", b.compute_fill())
trk = pkg.Tracker(rng.integers(0, 256, (H, W)).astype(np.uint8), 25, (45, 45), True)
print("tracker call", trk((40, 60))); trk.close()
print("done")
