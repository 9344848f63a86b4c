"""Chained tracking of 256 resident 1080p videos at other target widths (default window 4*ceil(sigma)+1):
us per step, algorithmic TFLOP/s and fraction of the FP32 peak, next to tw = 25 (the specialised kernels).
Usage: python tools/tw_sweep.py [--n 256] [--T 20] [tw ...]"""
import math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, pt_import
from tools import benchlib
pkg = pt_import.load()
H, W = bench.H, bench.W
args = sys.argv[1:]
n, T = 256, 20
while args and args[0].startswith("--"):
    if args[0] == "--n": n = int(args[1])
    if args[0] == "--T": T = int(args[1])
    args = args[2:]
tws = [float(x) for x in args] or [10, 15, 20, 25, 50]
dev = torch.device("cuda", 0)
peak = max(benchlib.fp32_peak(0, 0, 3), benchlib.fp32_peak(0, 1, 3))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pos = bench.orbit_positions(n, 0)
ring = bench.render_ring_device(torch, pos, T, dev)
for tw in tws:
    sigma = tw / (2 * math.sqrt(2 * math.log(2)))
    l = 4 * math.ceil(sigma * math.sqrt(2)) + 1
    ws = 4 * math.ceil(sigma) + 1
    b = pkg.TrackerBatch(n, (H, W), tw, (ws, ws), True)
    b.bind_device_frames(ring.data_ptr(), H * W, W); b.set_fill(128)
    b.set_guess(pos[0]); ij, _ = b.track_device(ring.data_ptr(), n * H * W, H * W, W, T)
    ok = bool(np.array_equal(ij, bench.truth_for_steps(pos, T)))
    ts = []
    for _ in range(7):
        benchlib.flush_l2(flush.data_ptr(), flush.numel(), b.stream)
        b.set_guess(pos[0]); b.track_device_async(ring.data_ptr(), n * H * W, H * W, W, 1)
        torch.cuda.synchronize()
        ts.append(benchlib.time_chain(pkg.lib, b, [(ring.data_ptr() + n * H * W, T - 1)], n * H * W, H * W, W, b.stream) * T / (T - 1))
        torch.cuda.synchronize()
    t = float(np.median(ts)) * 1e-3
    alg = bench.algorithmic_per_window(l, ws, ws)
    tf = n * T * alg["flops"] / t / 1e12
    print(f"tw={tw:5.1f} l={l:3d} ws={ws:3d} n={n} kernel={b.last_kernel:26s} {t/T*1e6:8.2f} us/step  "
          f"{tf:6.2f} TFLOP/s = {tf/peak:.3f} of {peak:.1f}  positions {'ok' if ok else 'differ from disk centres'}", flush=True)
    b.close()
