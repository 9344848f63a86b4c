"""Profiling aid: per-phase clock64 timings of dog_window45_argmax on the bench workload
(256 resident 1080p videos, chained steps).  Prints phase durations split by how many
CTAs shared the SM.  Usage: python tools/phase_timing.py [T]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import pt_import
pkg = pt_import.load()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n, H, W = bench.N_VIDEOS, bench.H, bench.W
dev = torch.device("cuda", 0)
pos = bench.orbit_positions(n, 0)
ring = bench.render_ring_device(torch, pos, 16 * ((T + 15) // 16), dev)
b = pkg.TrackerBatch(n, (H, W), bench.TW, (bench.WS, bench.WS), True)
b.bind_device_frames(ring.data_ptr(), H * W, W)
b.set_fill(128)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
names = ["stage", "row", "col", "reduce"]
for label in ("cold (L2 flushed)", "warm (same slots again)", "cold again"):
    if label.startswith("cold"):
        flush.fill_(1)
    dbg = torch.zeros((n, T, 6), dtype=torch.int64, device=dev)
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
    smid = d[:, 0, 0]
    cnt = np.bincount(smid, minlength=148)
    per_sm = cnt[smid]
    ghz = (d[:, -1, 1] - d[:, 1, 1]) / np.maximum(d[:, -1, 0] - d[:, 1, 0], 1)
    print(f"   in-kernel SM clock (clock64 / globaltimer): median {np.median(ghz):.3f} GHz, min {ghz.min():.3f}, max {ghz.max():.3f}")
    span = (d[:, -1, 5] - d[:, 0, 1])
    print(f"== {label}: kernel {ms*1e3:.1f} us = {ms*1e3/T:.2f} us/step; longest CTA span {span.max()} cycles "
          f"-> implied clock {span.max()/ms/1e6:.3f} GHz; SM CTA counts {np.bincount(cnt)}")
    per_cta = (d[:, -1, 5] - d[:, 2, 1]) / (T - 2)
    for k in (1, 2):
        mm = per_sm == k
        if mm.any():
            q = np.percentile(per_cta[mm], [0, 25, 50, 75, 100])
            print(f"   per-CTA mean frame period, SMs with {k}: min/25/50/75/max = " + "/".join(f"{x:.0f}" for x in q))
    for k in (1, 2):
        m = per_sm == k
        if not m.any():
            continue
        ph = np.diff(d[m][:, 2:, 1:], axis=2)
        frame = d[m][:, 3:, 1] - d[m][:, 2:-1, 1]
        first = np.diff(d[m][:, 0, 1:])
        print(f"   SMs with {k} CTA(s): frame period {frame.mean():.0f} cyc: "
              + " ".join(f"{nm} {ph[..., i].mean():.0f}" for i, nm in enumerate(names))
              + " | first frame: " + " ".join(f"{nm} {first[:, i].mean():.0f}" for i, nm in enumerate(names)))
