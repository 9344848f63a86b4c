"""Profiling aid: per-phase clock64 timings of dog_window45_argmax on the bench workload
(256 resident 1080p videos, chained steps).  Prints phase durations split by how many
CTAs shared the SM.  Usage: python tools/phase_timing.py [T]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the product library carries no probes: load the -DPT_PROBES build of the same sources
os.environ["PAWSOME_CUDA_LIB"] = os.path.join(ROOT, "pawsometracker.jl_b200", "libpawsome_cuda_probes.so")
import torch
import bench
import pt_import
pkg = pt_import.load()
import ctypes as _C
pkg.lib.pt_debug_window45_timing.restype = _C.c_int
pkg.lib.pt_debug_window45_timing.argtypes = [_C.c_void_p]
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n, H, W = bench.N_VIDEOS, bench.H, bench.W
dev = torch.device("cuda", 0)
pos = bench.orbit_positions(n, 0)
ring = bench.render_ring_device(torch, pos, 16 * ((T + 15) // 16), dev)
b = pkg.TrackerBatch(n, (H, W), bench.TW, (bench.WS, bench.WS), True)
NC = n
print("kernel:", b.kernel_name)
b.bind_device_frames(ring.data_ptr(), H * W, W)
b.set_fill(128)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
names = ["stage", "row", "col", "reduce"]
for label in ("cold (L2 flushed)", "warm (same slots again)", "cold again"):
    if label.startswith("cold"):
        flush.fill_(1)
    dbg = torch.zeros((NC, T, 6), dtype=torch.int64, device=dev)
    pkg.lib.pt_debug_window45_timing(dbg.data_ptr())
    b.set_guess(pos[0])
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(b.stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        e0.record()
        b.track_device_async(ring.data_ptr(), n * H * W, H * W, W, T)
        e1.record()
    torch.cuda.synchronize()
    ij, _ = b.read_track(T)
    pkg.lib.pt_debug_window45_timing(None)
    ms = e0.elapsed_time(e1)
    d = dbg.cpu().numpy()
    assert np.array_equal(ij, bench.truth_for_steps(pos, T))
    smid = d[:, :, 0] & 0xFF
    gt = d[:, :, 0] >> 8
    ph = np.diff(d[:, :, 1:], axis=2)                       # (n, T, 4): stage, row, col, reduce+publish
    wall = (gt.max() - gt.min()) / 1e3
    ghz = ph.sum() / 1.0                                     # placeholder, see below
    # windows in flight per SM at each window start: count windows on the same SM whose [start, end) covers it
    start = d[:, :, 1]; end = d[:, :, 5]
    print(f"== {label}: kernel {ms*1e3:.1f} us = {ms*1e3/T:.2f} us/step; first→last window start {wall:.1f} us")
    per_win = (end - start)
    print("   per-window cycles: mean %.0f  p50 %.0f  p90 %.0f  max %.0f | " % (per_win.mean(), np.median(per_win), np.percentile(per_win, 90), per_win.max())
          + " ".join(f"{nm} {ph[..., i].mean():.0f}" for i, nm in enumerate(names)))
    # hand-off latency: start of (v,t+1) minus end of (v,t) in wall-clock terms is not available (different SMs);
    # use globaltimer of consecutive frames of a video instead
    gap = np.diff(gt, axis=1) / 1e3                          # us between consecutive frame starts of a video
    print("   per-video frame period (wall us): mean %.2f  p50 %.2f  p90 %.2f  max %.2f" % (gap.mean(), np.median(gap), np.percentile(gap, 90), gap.max()))
    sm0 = smid[:, 0]
    cnt = np.bincount(sm0, minlength=148)
    per_cta = (end[:, -1] - start[:, 2]) / (T - 2)
    for k in sorted(set(cnt[sm0].tolist())):
        m = cnt[sm0] == k
        q = np.percentile(per_cta[m], [0, 25, 50, 75, 100])
        print(f"   windows on SMs hosting {k}: {m.sum():3d}; mean frame period min/25/50/75/max = " + "/".join(f"{x:.0f}" for x in q)
              + " | " + " ".join(f"{nm} {ph[m][:, 2:, i].mean():.0f}" for i, nm in enumerate(names)))

# ---- timeline of one SM that hosts two windows (last measured run)
try:
    sm0 = smid[:, 0]
    cnt = np.bincount(sm0, minlength=148)
    target_sm = int(np.flatnonzero(cnt == 2)[0])
    vs = np.flatnonzero(sm0 == target_sm)
    base = d[vs][:, :, 1].min()
    print(f"timeline on SM {target_sm}, videos {vs.tolist()} (cycles since first start; S=stage R=row C=col X=reduce):")
    for t in range(6, 10):
        for v in vs:
            st = d[v, t, 1:6] - base
            print(f"  v{v} t{t}: S {st[0]:7d}-{st[1]:7d} R -{st[2]:7d} C -{st[3]:7d} X -{st[4]:7d}")
except Exception as e:
    print("timeline failed:", e)
