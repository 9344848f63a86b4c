"""A single video (and a few) tracked from PAGE-LOCKED HOST frames handed to the chained kernels as if they were resident
(UVA pointer + strides): the cluster kernel then prefetches each step's region with one TMA tile copy over PCIe while
the previous step computes.  Compared with the frame-pointer-table path (global loads of the footprint after the guess
is known).  Usage: python tools/pinned_chain_timing.py [n ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, pt_import
pkg = pt_import.load()
H, W = bench.H, bench.W
ns = [int(x) for x in sys.argv[1:]] or [1, 4, 16]
for n in ns:
    T = 64 if n <= 32 else 16
    pos = bench.orbit_positions(n, 0)
    ring_dev = bench.render_ring_device(torch, pos, T, torch.device("cuda", 0))
    ring = torch.empty((T, n, H, W), dtype=torch.uint8, pin_memory=True)
    ring.copy_(ring_dev.cpu()); del ring_dev
    truth = bench.truth_for_steps(pos, T)
    b = pkg.TrackerBatch(n, (H, W), bench.TW, (bench.WS, bench.WS), True)
    b.set_fill([128] * n)
    variants = [("auto", 0, 1), ("C=4 tma", 4, 1), ("C=4 ldg", 4, 0), ("C=8 tma", 8, 1), ("C=2 tma", 2, 1), ("C=2 ldg", 2, 0), ("per-SM", 1, 0)]
    if n > 64: variants = [("auto", 0, 1), ("per-SM", 1, 0)]
    for name, C, bulk in variants:
        b.set_option("cluster", C); b.set_option("bulk", bulk)
        # (a) UVA pointer + strides
        b.bind_device_frames(ring.data_ptr(), H * W, W)
        ts = []
        for _ in range(5):
            b.set_guess(pos[0]); torch.cuda.synchronize()
            t0 = time.perf_counter(); ij, _ = b.track_device(ring.data_ptr(), n * H * W, H * W, W, T); ts.append(time.perf_counter() - t0)
        ok = bool(np.array_equal(ij, truth)); k1 = b.last_kernel
        # (b) frame-pointer table (pt_batch_track_host, zero-copy)
        steps = [[ring[t, v].numpy() for v in range(n)] for t in range(T)]
        tp = []
        for _ in range(5):
            b.set_guess(pos[0]); torch.cuda.synchronize()
            t0 = time.perf_counter(); ij2, _ = b.track_host(steps, mode="footprint"); tp.append(time.perf_counter() - t0)
        ok2 = bool(np.array_equal(ij2, truth)); k2 = b.last_kernel
        print(f"n={n:2d} {name:8s} UVA+strides: {min(ts)/T*1e6:6.2f} us/step {k1:24s} ok={ok} | pointer table: {min(tp)/T*1e6:6.2f} us/step {k2:24s} ok={ok2}", flush=True)
    b.close()
