import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import pt_import
pkg = pt_import.load()
f = np.full((480, 640), 128, np.uint8); f[200:225, 300:325] = 0
for rep in range(4):
    t0 = time.perf_counter(); b = pkg.TrackerBatch(1, f.shape, 25, (45, 45), True); t1 = time.perf_counter()
    b.set_frames([f]); t2 = time.perf_counter()
    b.compute_fill(); t3 = time.perf_counter()
    b.set_guess([[212, 312]]); o = b.track_host([[f]], mode="footprint"); t4 = time.perf_counter()
    o = b.track_host([[f]], mode="footprint"); t5 = time.perf_counter()
    b.close(); t6 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} ms  set_frames {1e3*(t2-t1):.2f}  compute_fill {1e3*(t3-t2):.2f}  first step {1e3*(t4-t3):.2f}  second step {1e3*(t5-t4):.3f}  close {1e3*(t6-t5):.2f}")
