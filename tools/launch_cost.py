"""Host-side cost of one chained launch call (pt_batch_track_device_async) per kernel family: the call's own time on
the host while the GPU is kept busy (so nothing waits), i.e. what sits between the CUDA event and the kernel when a
launch is timed from an idle GPU.  Usage: python tools/launch_cost.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, pt_import
pkg = pt_import.load()
H, W = bench.H, bench.W
dev = torch.device("cuda", 0)
T = 20
for n in (1, 32, 64, 256):
    pos = bench.orbit_positions(n, 0)
    ring = bench.render_ring_device(torch, pos, T, dev)
    b = pkg.TrackerBatch(n, (H, W), bench.TW, (bench.WS, bench.WS), True)
    b.bind_device_frames(ring.data_ptr(), H * W, W); b.set_fill([128] * n)
    b.set_guess(pos[0]); b.track_device(ring.data_ptr(), n * H * W, H * W, W, T)
    torch.cuda.synchronize()
    reps = 200
    t0 = time.perf_counter()
    for _ in range(reps):
        b.track_device_async(ring.data_ptr(), n * H * W, H * W, W, T)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"n={n:3d} {b.last_kernel:26s} host {1e6*(t1-t0)/reps:6.2f} us per call; GPU {1e6*(t2-t0)/reps:7.2f} us per call ({1e6*(t2-t0)/reps/T:5.2f} per step)")
    b.close(); del ring
