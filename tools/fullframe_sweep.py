"""Times the generic kernel on the full-frame 1080p DoG shape and the 1080p auto-detect window."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, pt_import
pkg = pt_import.load()
H, W = 1080, 1920
dev = torch.device("cuda", 0)
f = np.full((H, W), 128, np.uint8)
bench.render_frame_host(f, (700, 1234))
for n in (1, 16):
    b = pkg.TrackerBatch(n, (H, W), 25, (45, 45), True)
    b.set_frames([f] * n); b.set_fill(128)
    ext = torch.cuda.ExternalStream(b.stream, device=dev)
    for name, call in (("fullframe 1080x1920", lambda: b.rect_argmax(0, 0, 0, H, W)),):
        for _ in range(3): r = call()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            with torch.cuda.stream(ext):
                e0.record(); r = call(); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        if n == 1: print(f"{name}: {min(ts)*1e3:.1f} us  {H*W/min(ts)/1e3:.0f} MP/s  result {r[0]}")
    b.set_window((270, 480))
    g = np.tile([540, 960], (n, 1))
    for _ in range(3): b.step(g)
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):
            e0.record(); o = b.step(g); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"autodetect 271x481 x{n}: {min(ts)*1e3:.1f} us ({min(ts)*1e3/n:.1f} us/window) result {o[0][0]}")
    b.close()
