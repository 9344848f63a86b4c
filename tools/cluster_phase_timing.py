"""Phase timestamps (clock64, -DPT_PROBES build) of the lone-window cluster kernel on the bench geometry:
per step of rank 0 of every cluster: stage (wait for the prefetched region + u8→f32 conversion, up to the first CTA
barrier), row pass, column pass + argmax exchange (up to the cluster barrier), publish.
Usage: python tools/cluster_phase_timing.py [n ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PAWSOME_CUDA_LIB"] = os.path.join(ROOT, "pawsometracker.jl_b200", "libpawsome_cuda_probes.so")
import ctypes as C
import torch, bench, pt_import
pkg = pt_import.load()
pkg.lib.pt_debug_window45_timing.restype = C.c_int
pkg.lib.pt_debug_window45_timing.argtypes = [C.c_void_p]
H, W, T = bench.H, bench.W, 40
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in [int(x) for x in sys.argv[1:]] or [1, 32]:
    pos = bench.orbit_positions(n, 0)
    ring = bench.render_ring_device(torch, pos, T, dev)
    b = pkg.TrackerBatch(n, (H, W), bench.TW, (bench.WS, bench.WS), True)
    b.bind_device_frames(ring.data_ptr(), H * W, W); b.set_fill(128)
    ext = torch.cuda.ExternalStream(b.stream, device=dev)
    for Cc in (1, 2, 4, 8):
        for bulk in ((0,) if Cc == 1 else (0, 1)):
            b.set_option("cluster", Cc); b.set_option("bulk", bulk)
            for rep in range(3):
                flush.fill_(1)
                dbg = torch.zeros((n, T, 6), dtype=torch.int64, device=dev)
                pkg.lib.pt_debug_window45_timing(dbg.data_ptr())
                b.set_guess(pos[0]); b.track_device(ring.data_ptr(), n * H * W, H * W, W, 3)      # warm-up, uploads the guess
                dbg.zero_(); b.set_guess(pos[0]); b.track_device(ring.data_ptr(), n * H * W, H * W, W, 1)
                dbg.zero_(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(ext):
                    e0.record(); b.track_device_async(ring.data_ptr() + H * W * n, n * H * W, H * W, W, T - 1); e1.record()
                torch.cuda.synchronize()
                pkg.lib.pt_debug_window45_timing(None)
            ms = e0.elapsed_time(e1)
            d = dbg.cpu().numpy()[:, :T - 1]
            ph = np.diff(d[:, :, 1:], axis=2)                  # stage, row, col+exchange, publish
            period = np.diff(d[:, :, 1], axis=1)               # start-to-start of consecutive steps
            gt = (d[:, :, 0] >> 8)
            wall = (gt[:, -1] - gt[:, 0]).mean() / 1e3 / (T - 2)
            print(f"n={n:3d} C={Cc} bulk={bulk} {b.last_kernel:24s} {ms*1e3/(T-1):5.2f} us/step by events, {wall:5.2f} us/step start-to-start | "
                  f"cycles/step {period[:, 2:].mean():6.0f}: stage {ph[:, 2:, 0].mean():5.0f} row {ph[:, 2:, 1].mean():5.0f} "
                  f"col+xchg {ph[:, 2:, 2].mean():5.0f} publish {ph[:, 2:, 3].mean():5.0f}", flush=True)
    b.close(); del ring
