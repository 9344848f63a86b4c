"""Kernel-only timing of the l=65 rectangle kernel (dog_rect45_march) on the full-frame 1080p DoG shape
(n frames per launch, launches enqueued back to back without read-back, CUDA events on the batch stream)
and on the 1080p auto-detect window.  Usage: python tools/rect_timing.py [n ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, pt_import
pkg = pt_import.load()
H, W = 1080, 1920
dev = torch.device("cuda", 0)
ns = [int(x) for x in sys.argv[1:]] or [1, 4, 16]
f = np.full((H, W), 128, np.uint8)
bench.render_frame_host(f, (700, 1234))
alg = bench.algorithmic_per_window(65, H, W)
for n in ns:
    b = pkg.TrackerBatch(n, (H, W), 25, (45, 45), True)
    b.set_frames([f] * n); b.set_fill(128)
    ext = torch.cuda.ExternalStream(b.stream, device=dev)
    ij, raw, resp = b.rect_argmax_all(0, 0, H, W)
    ok = bool((ij == [700, 1234]).all())
    reps = 20
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):
            e0.record()
            for _ in range(reps):
                b.rect_argmax_all(0, 0, H, W, readback=False)
            e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    t = best * 1e-3
    print(f"fullframe 1080x1920 x{n}: {best*1e3/n:.1f} us/frame  {n*H*W/t/1e9:.1f} GP/s  "
          f"{n*alg['flops']/t/1e12:.1f} TFLOP/s alg  correct={ok}")
    b.set_window((270, 480))
    g = np.tile([540, 960], (n, 1)).astype(np.int32)
    b.step(g)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):
            e0.record()
            for _ in range(reps):
                b.rect_argmax_all(405 - 1, 720 - 1, 271, 481, readback=False)
            e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    print(f"autodetect 271x481 x{n}: {best*1e3/n:.1f} us/window")
    b.close()
