"""Times BASELINE config 4 shapes (3840x2160, light target, target_width=100 → l=245) on the generic kernel."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, pt_import
pkg = pt_import.load()
H, W = 2160, 3840
dev = torch.device("cuda", 0)
f = np.full((H, W), 128, np.uint8)
yy, xx = np.ogrid[0:H, 0:W]
f[(yy - 1000) ** 2 + (xx - 2000) ** 2 <= 2500] = 255
# wide: 0 = 32-column kernel, 1 = fused 64-column kernel, 2 = 64-column kernel in two phases, 3 = automatic choice
cases = [(n, ws, wide) for n in (1, 4, 16, 64) for ws in (173, 401) for wide in (3, 2, 1, 0)]
if len(sys.argv) == 4:                      # one case: n ws wide   (for ncu)
    cases = [tuple(int(x) for x in sys.argv[1:4])]
for n, ws, wide in cases:
    if True:
        b = pkg.TrackerBatch(n, (H, W), 100, (ws, ws), False)
        b.set_option("wide", min(wide, 1)); b.set_option("two_phase", {0: 0, 1: 0, 2: 2, 3: 1}[wide])
        b.set_frames([f] * n); b.set_fill(128)
        ext = torch.cuda.ExternalStream(b.stream, device=dev)
        g = np.tile([1010, 1990], (n, 1))
        b.set_guess(g)
        for _ in range(3): o = b.step(None); b.set_guess(g)
        ts = []
        for _ in range(8):
            b.set_guess(g)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            with torch.cuda.stream(ext):
                e0.record(); o = b.step(None); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        l = 245; w = 122; wr = 2 * (ws // 2) + 1
        mac = 2 * l * wr * (wr + 2 * w + wr)
        t = min(ts) * 1e-3
        print(f"{b.last_kernel} n={n} ws={ws}: {t*1e6:.1f} us ({t*1e6/n:.1f} us/window), result {o[0][0]}, "
              f"{n*2*mac/t/1e12:.1f} TFLOP/s algorithmic")
        b.close()
