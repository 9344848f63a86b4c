"""GPU parity tests, round 2 (second part): target widths below 25 and windows below 45x45 through the per-window
kernels (dog_window45_argmax / _rot / _cluster<C>).  Those kernels are compiled for l = 65 and 45x45 outputs; a shorter
kernel runs zero-padded (0·x changes no sum), a smaller window masks the surplus outputs and skips footprint rows and
column groups without work.  Every variant against the oracle loop: positions exact, responses within RTOL·max|R|;
and against the generic streaming kernel (the path these geometries took before).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5
# a window that sees nothing but the border fill: the GPU path gives exactly 0, the f64 oracle its rounding noise (1e-16)
BLANK = 1e-12


def default_window(tw):
    import math
    return 4 * math.ceil(tw / (2 * math.sqrt(2 * math.log(2)))) + 1          # src/PawsomeTracker.jl:64-68


def oracle_chain(oracle, frames, tw, darker, ws, start_guess, fill):
    g = tuple(int(x) for x in start_guess)
    pos, resp, mx, near = [], [], [], 0
    for f in frames:
        r = oracle.step(f, fill, tw, darker, ws, g, dense=True)
        near += r.near_tie(RTOL)
        g = (r.i, r.j)
        pos.append(g); resp.append(r.resp); mx.append(r.maxabs)
    return np.array(pos), np.array(resp), np.array(mx), near


def make_case(synth, H, W, n, T, tw, darker, seed):
    vids = [synth.make_video(H=H, W=W, target_width=tw, start_ij=(H // 2, W // 2), seconds=10.0, fps=24.0, seed=seed + s,
                             darker_target=darker) for s in range(n)]
    frames = np.stack([np.stack([v.frame(t) for v in vids]) for t in range(T)])
    start = np.tile([H // 2, W // 2], (n, 1)).astype(np.int32)
    if n > 2:
        start[1] = (2, 3)                 # the window hangs over the top-left corner
        start[2] = (H + 7, W - 1)         # a guess outside the frame
    return frames, start


@pytest.mark.parametrize("tw,ws,darker", [
    (10, None, True), (15, None, True), (20, None, False), (22.5, None, True), (12, (21, 33), True),
    (25, (31, 45), True), (25, (45, 17), False), (8, (5, 9), True), (25, (1, 1), True),
])
def test_small_geometries_through_the_window_kernels(gpu_pkg, oracle, synth, tw, ws, darker):
    """Chained steps of 7 videos (one SM each) and of the same videos through the cluster kernel (2, 4, 8 CTAs per
    window, TMA and global-load staging), against the oracle loop and the generic kernel."""
    import torch
    ws = ws or (default_window(tw),) * 2
    n, T, H, W = 7, 6, 160, 208
    frames, start = make_case(synth, H, W, n, T, tw, darker, 500)
    dev = torch.from_numpy(frames).cuda()
    results = {}
    with gpu_pkg.TrackerBatch(n, (H, W), tw, ws, darker) as b:
        b.bind_device_frames(dev.data_ptr(), H * W, W)
        fills = b.compute_fill()
        for name, opts in [("per-SM", {"cluster": 1}), ("C2", {"cluster": 2, "bulk": 1}), ("C4", {"cluster": 4, "bulk": 0}),
                           ("C8", {"cluster": 8, "bulk": 1}), ("generic", {"window45": 0})]:
            for k, v in opts.items():
                b.set_option(k, v)
            b.set_guess(start)
            results[name] = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
            if name == "per-SM":
                assert b.last_kernel == "dog_window45_argmax"
            elif name.startswith("C"):
                assert b.last_kernel == f"dog_window45_cluster<{name[1]}>"
            else:
                assert b.last_kernel.startswith(("dog_rect_argmax", "dog_rows_wide"))
    ij0, r0 = results["per-SM"]
    for name in ("C2", "C4", "C8"):
        np.testing.assert_array_equal(results[name][0], ij0)
        np.testing.assert_array_equal(results[name][1], r0)           # same per-output operation order: bit-identical
    for v in range(n):
        pos, resp, mx, near = oracle_chain(oracle, [frames[t, v] for t in range(T)], tw, darker, ws, start[v], int(fills[v]))
        if near == 0:
            np.testing.assert_array_equal(ij0[:, v], pos)
            np.testing.assert_array_equal(results["generic"][0][:, v], pos)
        assert np.all(np.abs(r0[:, v] - resp) <= np.maximum(RTOL * mx, BLANK))


@pytest.mark.parametrize("tw,dtype", [(10, np.uint8), (20, np.uint8), (15, np.float32)])
def test_small_geometries_in_large_batches(gpu_pkg, oracle, tw, dtype):
    """More windows than SMs: the static split (two windows per SM) and the rotating-slot kernel, 11 chained steps of
    random-noise frames (every response value informative) — identical to each other, a sample against the oracle."""
    import torch
    ws = (default_window(tw),) * 2
    n, T, H, W = 230, 11, 96, 128
    rng = np.random.default_rng(int(tw))
    base = rng.integers(0, 256, (T, n, H, W)).astype(np.uint8)
    frames = base if dtype is np.uint8 else base.astype(np.float32) / np.float32(255.0)
    dev = torch.from_numpy(frames).cuda()
    start = np.stack([rng.integers(1, H + 1, n), rng.integers(1, W + 1, n)], axis=-1)
    with gpu_pkg.TrackerBatch(n, (H, W), tw, ws, True, dtype=dtype) as b:
        b.bind_device_frames(dev.data_ptr(), H * W, W)
        fills = b.compute_fill()
        b.set_option("rot", 2)
        b.set_guess(start)
        ij_rot, r_rot = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
        assert b.last_kernel == "dog_window45_rot"
        b.set_option("rot", 0)
        b.set_guess(start)
        ij_st, r_st = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
        assert b.last_kernel == "dog_window45_argmax"
    np.testing.assert_array_equal(ij_rot, ij_st)
    np.testing.assert_array_equal(r_rot, r_st)
    for v in (0, 1, 77, 148, n - 1):
        pos, resp, mx, near = oracle_chain(oracle, [base[t, v] for t in range(T)], tw, True, ws, start[v], int(fills[v]))
        if near == 0:
            np.testing.assert_array_equal(ij_st[:, v], pos)
        assert np.all(np.abs(r_st[:, v] - resp) <= np.maximum(RTOL * mx, BLANK))


def test_small_geometry_host_paths_and_single_tracker(gpu_pkg, oracle, synth):
    """Tracker(img, 15, default window) — the per-frame call on a host frame, the resident step, the response map
    (marching kernel with zero-padded taps) and track() on pageable and page-locked frames."""
    tw = 15
    ws = (default_window(tw),) * 2
    vid = synth.make_video(H=240, W=320, target_width=tw, start_ij=(120, 160), seconds=4.0, fps=24.0, seed=3)
    frames = np.stack([vid.frame(k) for k in range(40)])
    fill = oracle.mode(frames[0])
    trk = gpu_pkg.Tracker(frames[0], tw, ws, True)
    try:
        g = (118, 163)
        ref = oracle.step(frames[0], fill, tw, True, ws, g, dense=True, want_map=True)
        assert trk(g) == (ref.i, ref.j)
        assert abs(trk.last_response - ref.resp) <= RTOL * ref.maxabs
        assert trk.step_resident(g) == (ref.i, ref.j)
        rmap = trk.response_map(g)
        assert np.abs(rmap - ref.R).max() <= RTOL * ref.maxabs
    finally:
        trk.close()
    pos, _, _, near = oracle_chain(oracle, list(frames), tw, True, ws, (120, 160), fill)
    assert near == 0
    for pinned in (False, True):
        pin = None
        src = frames
        if pinned:
            pin = gpu_pkg.PinnedArray(frames.shape, np.uint8)
            pin.array[...] = frames
            src = pin.array
        av = gpu_pkg.ArrayVideo(src, fps=24.0)
        ts, ij = gpu_pkg.track(av, stop=len(frames) / 24.0, target_width=tw, start_location=gpu_pkg.CartesianIndex(120, 160),
                               darker_target=True, fps=24)
        np.testing.assert_array_equal(np.asarray(ij), pos)
        if pin is not None:
            del av
            pin.close()


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_two_phase_wide_path_in_batches_and_host_lanes(gpu_pkg, oracle, synth, dtype):
    """dog_rows_wide + dog_cols_wide (row-pass intermediate in global memory, one slice per window / per host lane):
    a batch of chained resident steps and the pageable host-frame path (crops through the lanes) give the fused
    kernel's results bit for bit and the oracle's positions; several row chunks per strip, windows over the frame edge."""
    import torch
    tw, ws, darker = 40, (73, 150), False
    n, T, H, W = 5, 4, 220, 300
    frames, start = make_case(synth, H, W, n, T, tw, darker, 900)
    fr = frames if dtype is np.uint8 else frames.astype(np.float32) / np.float32(255.0)
    dev = torch.from_numpy(fr).cuda()
    out = {}
    with gpu_pkg.TrackerBatch(n, (H, W), tw, ws, darker, dtype=dtype) as b:
        b.bind_device_frames(dev.data_ptr(), H * W, W)
        fills = b.compute_fill()
        for tp in (0, 2, 1):
            b.set_option("two_phase", tp)
            b.set_guess(start)
            out["dev", tp] = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
            assert b.last_kernel.startswith("dog_rect_argmax_wide" if tp == 0 else "dog_rows_wide"), b.last_kernel
            b.set_guess(start)
            out["host", tp] = b.track_host([[fr[t, v] for v in range(n)] for t in range(T)], mode="footprint")
        # page-locked frames: read in place by the streaming kernels, all T steps enqueued at once (no crops, no lanes)
        pin = gpu_pkg.PinnedArray(fr.shape, fr.dtype)
        pin.array[...] = fr
        b.set_option("two_phase", 1)
        b.set_guess(start)
        lc = b.launch_count
        out["pinned", 1] = b.track_host([[pin.array[t, v] for v in range(n)] for t in range(T)], mode="footprint")
        assert b.launch_count - lc == 3 * T, "pinned frames: footprint gather + two filter launches per step, no per-lane launches"
        b.set_option("crop_gather", 0)                           # the same with the filter kernels reading the host frames in place
        b.set_guess(start)
        out["pinned in place", 1] = b.track_host([[pin.array[t, v] for v in range(n)] for t in range(T)], mode="footprint")
        b.set_option("crop_gather", 1)
        nxt, _ = b.step(None)                                    # the chain state was left on the device
        pin.close()
    for key, (ij, r) in out.items():
        np.testing.assert_array_equal(ij, out["dev", 0][0], err_msg=str(key))
        np.testing.assert_array_equal(r, out["dev", 0][1], err_msg=str(key))
    ij0, r0 = out["dev", 0]
    for v in range(n):
        pos, resp, mx, near = oracle_chain(oracle, [frames[t, v] for t in range(T)], tw, darker, ws, start[v], int(fills[v]))
        if near == 0:
            np.testing.assert_array_equal(ij0[:, v], pos)
        assert np.all(np.abs(r0[:, v] - resp) <= np.maximum(RTOL * mx, BLANK))


@pytest.mark.parametrize("tw,ws", [(100, (150, 70)), (70, (121, 121))])
def test_two_team_column_kernel(gpu_pkg, oracle, tw, ws):
    """dog_cols_wide2 (two teams of 8 warps per CTA on two consecutive output batches, for kernels whose ring fills an SM)
    against the one-team kernel for several chunk heights: bit-identical response maps and results; map vs the oracle."""
    darker = False
    l = oracle.kernel_len(tw)
    rng = np.random.default_rng(int(tw))
    H, W = ws[0] + 40, ws[1] + 64
    f8 = rng.integers(90, 170, (H, W)).astype(np.uint8)
    yy, xx = np.ogrid[0:H, 0:W]
    f8[(yy - H // 2 - 5) ** 2 + (xx - W // 2 + 7) ** 2 <= (int(tw) // 2) ** 2] = 250
    guess = (H // 2, W // 2)
    fill = oracle.mode(f8)
    ref = oracle.step(f8, fill, tw, darker, ws, guess, dense=False, want_map=True)
    maps, outs = {}, {}
    for ch in (0, 64, 96, 160):
        for teams in (0, 1):
            trk = gpu_pkg.Tracker(f8, tw, ws, darker)
            try:
                trk.set_option("two_phase", 2); trk.set_option("cols_ch", ch); trk.set_option("cols_teams", teams)
                outs[ch, teams] = (trk.step_resident(guess), trk.last_response)
                maps[ch, teams] = trk.response_map(guess)
                assert trk._batch.last_kernel.startswith("dog_rows_wide")
            finally:
                trk.close()
    base = maps[0, 1]
    assert np.abs(base.astype(np.float64) - ref.R).max() <= RTOL * ref.maxabs
    for k, m in maps.items():
        np.testing.assert_array_equal(m, base, err_msg=str(k))
        assert outs[k] == outs[0, 1], k
    if not ref.near_tie(RTOL):
        assert outs[0, 1][0] == (ref.i, ref.j)


def test_two_phase_chain_runs_ahead_safely(gpu_pkg, synth):
    """The row and column kernels of consecutive steps are programmatic dependent launches of one another: a kernel may
    start before its predecessor has finished and must read nothing the predecessor writes (the published guess above
    all) before its griddepcontrol.wait.  A long chain over page-locked host frames — slow row kernels, the widest window
    for a kernel to run ahead in — must equal the chain over resident frames and the step-by-step calls."""
    import torch
    tw, ws, darker = 55, (97, 97), True
    n, T, H, W = 6, 14, 178, 88
    frames, start = make_case(synth, H, W, n, T, tw, darker, 1234)
    dev = torch.from_numpy(frames).cuda()
    pin = gpu_pkg.PinnedArray(frames.shape, frames.dtype)
    pin.array[...] = frames
    with gpu_pkg.TrackerBatch(n, (H, W), tw, ws, darker) as b:
        b.bind_device_frames(dev.data_ptr(), H * W, W)
        b.compute_fill()
        b.set_guess(start)
        ij_res, r_res = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
        assert b.last_kernel.startswith("dog_rows_wide")
        for rep in range(3):
            b.set_guess(start)
            ij_pin, r_pin = b.track_host([[pin.array[t, v] for v in range(n)] for t in range(T)], mode="footprint")
            np.testing.assert_array_equal(ij_pin, ij_res)
            np.testing.assert_array_equal(r_pin, r_res)
        b.set_guess(start)
        per_step = []
        for t in range(T):
            b.bind_device_frames(dev.data_ptr() + t * n * H * W, H * W, W)
            o, rr = b.step(None)
            per_step.append(o.copy())
        np.testing.assert_array_equal(np.stack(per_step), ij_res)
    pin.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_track_over_one_array_of_frames(gpu_pkg, oracle, synth, pinned):
    """track() on a video held as ONE (T, H, W) array (pageable or page-locked): whole blocks of frame addresses go to the
    library in one call each; with start > 0 and fps resampling the positions equal the oracle loop on the selected frames."""
    tw = 25
    vid = synth.make_video(H=200, W=256, target_width=tw, start_ij=(100, 128), seconds=6.0, fps=24.0, seed=9)
    frames = np.stack([vid.frame(k) for k in range(144)])
    pin = None
    if pinned:
        pin = gpu_pkg.PinnedArray(frames.shape, np.uint8)
        pin.array[...] = frames
    src = pin.array if pinned else frames
    start, stop, fps = 0.5, 5.5, 12.0
    ts, ij = gpu_pkg.track(gpu_pkg.ArrayVideo(src, fps=24.0), start=start, stop=stop, target_width=tw,
                           start_location=gpu_pkg.CartesianIndex(*[int(x) for x in vid.traj[12]]), darker_target=True, fps=fps)
    n = int(round(fps * (stop - start)))
    sel = [int(np.floor((start + k / fps) * 24.0 + 0.5)) for k in range(n)]
    sel = [k for k in sel if k < len(frames)]
    assert len(ij) == len(sel)
    fill = oracle.mode(frames[sel[0]])
    g = tuple(int(x) for x in vid.traj[12])
    for k, fi in enumerate(sel):
        r = oracle.step(frames[fi], fill, tw, True, (45, 45), g, dense=True)
        assert tuple(ij[k]) == (r.i, r.j), k
        g = (r.i, r.j)
    if pin is not None:
        pin.close()
