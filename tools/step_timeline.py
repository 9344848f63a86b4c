"""Where the fixed cost of a chained launch goes: globaltimer stamps at every (window, step) start of
dog_window45_rot on the bench workload → per-step median start time, spread, and the last steps' tail.
Usage: python tools/step_timeline.py [T]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the product library carries no probes: load the -DPT_PROBES build of the same sources
os.environ["PAWSOME_CUDA_LIB"] = os.path.join(ROOT, "pawsometracker.jl_b200", "libpawsome_cuda_probes.so")
import torch, bench, pt_import
pkg = pt_import.load()
import ctypes as _C
pkg.lib.pt_debug_window45_timing.restype = _C.c_int
pkg.lib.pt_debug_window45_timing.argtypes = [_C.c_void_p]
T = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n, H, W = bench.N_VIDEOS, bench.H, bench.W
dev = torch.device("cuda", 0)
pos = bench.orbit_positions(n, 0)
ring = bench.render_ring_device(torch, pos, 16 * ((T + 15) // 16), dev)
b = pkg.TrackerBatch(n, (H, W), bench.TW, (bench.WS, bench.WS), True)
b.bind_device_frames(ring.data_ptr(), H * W, W); b.set_fill(128)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ext = torch.cuda.ExternalStream(b.stream, device=dev)
for rep in range(3):
    flush.fill_(1)
    dbg = torch.zeros((n, T, 6), dtype=torch.int64, device=dev)
    pkg.lib.pt_debug_window45_timing(dbg.data_ptr())
    b.set_guess(pos[0]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        e0.record(); b.track_device_async(ring.data_ptr(), n * H * W, H * W, W, T); e1.record()
    torch.cuda.synchronize()
    pkg.lib.pt_debug_window45_timing(None)
ms = e0.elapsed_time(e1)
d = dbg.cpu().numpy()
gt = (d[:, :, 0] >> 8).astype(np.float64) / 1e3          # us
t0 = gt.min()
st = gt - t0
print(f"kernel {ms*1e3:.1f} us ({b.last_kernel}); first window start = 0, per step: median / min / max start (us), median period")
prev = None
for t in range(T):
    med = np.median(st[:, t])
    print(f"  t={t:2d}: {med:7.1f} / {st[:, t].min():7.1f} / {st[:, t].max():7.1f}   period {med - prev if prev is not None else 0:5.2f}")
    prev = med
print(f"last step starts: median {np.median(st[:, -1]):.1f}, max {st[:, -1].max():.1f}; kernel end (event) - first start unknown; "
      f"event time - last median start = {ms*1e3 - np.median(st[:, -1]):.1f} us (includes launch latency before the first start)")
if os.environ.get("PT_TIMELINE_DUMP"):
    os.makedirs(os.path.dirname(os.environ["PT_TIMELINE_DUMP"]) or ".", exist_ok=True)
    np.save(os.environ["PT_TIMELINE_DUMP"], d)
