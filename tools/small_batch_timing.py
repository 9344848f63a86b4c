"""Chained tracking of small batches (n <= #SMs): frames resident in HBM, T steps per launch, CUDA events.
A lone window needs ≈ 6.1 us per step (12 K cycles) whatever n ≤ #SMs is.  Usage: python tools/small_batch_timing.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, pt_import
pkg = pt_import.load()
H, W, T = bench.H, bench.W, 32
dev = torch.device("cuda", 0)
for n in [int(x) for x in sys.argv[1:]] or (1, 32, 100, 148):
    pos = bench.orbit_positions(n, 0)
    ring = bench.render_ring_device(torch, pos, T, dev)
    b = pkg.TrackerBatch(n, (H, W), bench.TW, (bench.WS, bench.WS), True)
    b.bind_device_frames(ring.data_ptr(), H * W, W); b.set_fill(128)
    ext = torch.cuda.ExternalStream(b.stream, device=dev)
    b.set_guess(pos[0]); ij, _ = b.track_device(ring.data_ptr(), n * H * W, H * W, W, T)
    ok = bool(np.array_equal(ij, bench.truth_for_steps(pos, T)))
    best = 1e9
    for _ in range(10):
        b.set_guess(pos[0]); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record(); b.track_device_async(ring.data_ptr(), n * H * W, H * W, W, T); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"n={n:3d}: {best*1e3/T:6.2f} us per step ({n*T/best:8.1f} k frames/s)  kernel {b.last_kernel}  correct={ok}")
    b.close(); del ring
