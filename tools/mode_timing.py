"""Times pt_batch_compute_fill (mode of every frame, src/PawsomeTracker.jl:47) on 64 1080p u8 frames:
flat background + disk, mild sensor-like noise (+-3), uniform noise 0..255.  CUDA events on the batch stream,
the call includes the read-back of the fills."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pt_import
pkg = pt_import.load()
H, W, n = 1080, 1920, 64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
flat = torch.full((n, H, W), 128, dtype=torch.uint8, device=dev); flat[:, 500:525, 900:925] = 0
mild = (128 + torch.randint(-3, 4, (n, H, W), device=dev, generator=g)).to(torch.uint8)
noise = torch.randint(0, 256, (n, H, W), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
torch.cuda.synchronize()
b = pkg.TrackerBatch(n, (H, W), 25, (45, 45), True)
ext = torch.cuda.ExternalStream(b.stream, device=dev)
for name, fr in (("flat+disk", flat), ("mild noise +-3", mild), ("uniform noise", noise)):
    b.bind_device_frames(fr.data_ptr(), H * W, W)
    fills = b.compute_fill()
    ref = [int(torch.bincount(fr[v].flatten().int(), minlength=256).argmax()) for v in range(3)]
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):
            e0.record(); b.compute_fill(); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = min(ts) * 1e-3
    print(f"{name:16s}: {t*1e6:7.1f} us for {n} frames = {n*H*W/t/1e9:7.1f} GB/s   fills {list(fills[:3])} (bincount argmax {ref})")
b.close()
