"""Chained tracking of small batches (n <= #SMs): frames resident in HBM, T steps per launch, CUDA events, L2 flushed
before every launch.  Sweeps the lone-window options: the per-SM kernel (cluster=1), the cluster kernel with 2 / 4 / 8
CTAs per window and its two staging modes (bulk 0 = global loads + L2 prefetch, 1 = one TMA tile copy per step),
and the automatic choice.
Usage: python tools/small_batch_timing.py [--T 20,100] [n ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, pt_import
from tools import benchlib
pkg = pt_import.load()
H, W = bench.H, bench.W
args = sys.argv[1:]
Ts = [20, 100]
if args and args[0] == "--T":
    Ts = [int(x) for x in args[1].split(",")]
    args = args[2:]
ns = [int(x) for x in args] or [1, 8, 18, 32, 37, 64, 74, 100, 128, 148]
dev = torch.device("cuda", 0)
sms = torch.cuda.get_device_properties(dev).multi_processor_count
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
Tmax = max(Ts)
for n in ns:
    pos = bench.orbit_positions(n, 0)
    ring = bench.render_ring_device(torch, pos, Tmax, dev)
    b = pkg.TrackerBatch(n, (H, W), bench.TW, (bench.WS, bench.WS), True)
    b.bind_device_frames(ring.data_ptr(), H * W, W); b.set_fill(128)
    ext = torch.cuda.ExternalStream(b.stream, device=dev)
    variants = [("auto", 0, 1), ("per-SM", 1, 0)]
    for C in (2, 4, 8):
        if n * C <= 4 * sms:
            variants += [(f"C={C} bulk=1", C, 1), (f"C={C} bulk=0", C, 0)]
    for name, C, bulk in variants:
        b.set_option("cluster", C); b.set_option("bulk", bulk)
        row = []
        for T in Ts:
            b.set_guess(pos[0]); ij, _ = b.track_device(ring.data_ptr(), n * H * W, H * W, W, T)
            ok = bool(np.array_equal(ij, bench.truth_for_steps(pos, T)))
            ts = []
            for _ in range(7):
                benchlib.flush_l2(flush.data_ptr(), flush.numel(), b.stream)
                # 3 untimed warm-up steps (they also upload the guess), then T timed steps — bench.py's protocol
                b.set_guess(pos[(Tmax - 3) % bench.PERIOD]); b.track_device_async(ring.data_ptr() + (Tmax - 3) * n * H * W, n * H * W, H * W, W, 3)
                b.set_guess(pos[0]); b.track_device_async(ring.data_ptr(), n * H * W, H * W, W, 1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(ext):
                    e0.record(); b.track_device_async(ring.data_ptr() + n * H * W, n * H * W, H * W, W, T - 1); e1.record()
                torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * T / (T - 1))
            t = float(np.median(ts))
            row.append(f"T={T}: {t*1e3/T:6.2f} us/step ({n*T/t/1e3:7.2f} M frames/s){'' if ok else ' WRONG'}")
        print(f"n={n:3d} {name:12s} {b.last_kernel:24s} " + "   ".join(row), flush=True)
    b.close(); del ring
