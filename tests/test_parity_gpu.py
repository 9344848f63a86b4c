"""GPU parity tests: the CUDA path (through the C ABI of libpawsome_cuda.so)
against the CPU oracle on identical synthetic frames.

Bars (BASELINE.json north_star / SURVEY §8c):
  * argmax position: EXACT, except frames the oracle flags as near-ties
    (top-2 gap < RTOL·max|R|), which are reported and skipped;
  * response: |R_gpu − R_oracle| ≤ RTOL·max|R_oracle| with RTOL = 1e-5 (FP32
    separable vs the reference-order dense Float64 sum).
"""
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def disk_frame(H, W, cy, cx, r, val=0, bg=128):
    f = np.full((H, W), bg, np.uint8)
    yy, xx = np.ogrid[0:H, 0:W]
    f[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = val
    return f


def check_step(pkg, oracle, frame, tw, darker, ws, guess, dense=True, check_map=True, options=None):
    trk = pkg.Tracker(frame, tw, ws, darker)
    try:
        for k_, v_ in (options or {}).items():
            trk.set_option(k_, v_)
        ofill = oracle.mode(frame)
        assert trk.fillvalue == ofill
        ref = oracle.step(frame, ofill, tw, darker, ws, guess, dense=dense, want_map=check_map)
        got_host = trk(guess)                      # footprint-streaming path
        resp_host = trk.last_response
        got_res = trk.step_resident(guess)         # frame resident in HBM
        resp_res = trk.last_response
        assert got_host == got_res, "footprint and resident paths disagree"
        assert resp_host == resp_res
        tol = RTOL * ref.maxabs
        assert abs(resp_res - ref.resp) <= tol, (resp_res, ref.resp, ref.maxabs)
        if check_map:
            rmap = trk.response_map(guess)
            assert rmap.shape == ref.R.shape
            err = np.abs(rmap.astype(np.float64) - ref.R).max()
            assert err <= tol, f"response map error {err / ref.maxabs:.3e} of max|R|"
        if ref.near_tie(RTOL) and ref.resp != ref.second:
            warnings.warn(f"documented near-tie: top-2 gap {(ref.resp - ref.second) / ref.maxabs:.2e} of max|R|; "
                          "argmax not compared")
        else:
            # incl. EXACT ties (resp == second): findmax's first-in-column-major rule decides, in both implementations
            assert got_res == (ref.i, ref.j), (got_res, (ref.i, ref.j))
        return ref
    finally:
        trk.close()


# ---------------------------------------------------------------------------
def test_golden_vectors(gpu_pkg, oracle):
    z = np.load(os.path.join(GOLDEN, "oracle_cases.npz"))
    for c in range(int(z["n"])):
        f = z[f"frame{c}"]
        tw, darker, wsr, wsc, gi, gj, fill = [z[f"par{c}"][k] for k in range(7)]
        exp = z[f"res{c}"]
        trk = gpu_pkg.Tracker(f, float(tw), (int(wsr), int(wsc)), bool(darker))
        try:
            assert trk.fillvalue == int(fill)
            assert trk((int(gi), int(gj))) == (int(exp[0]), int(exp[1]))
            assert abs(trk.last_response - exp[4]) <= RTOL * exp[6]
            rmap = trk.response_map((int(gi), int(gj)))
            assert np.abs(rmap - z[f"map{c}"]).max() <= RTOL * exp[6]
        finally:
            trk.close()


@pytest.mark.parametrize("cy,cx,guess", [
    (240, 320, (236, 327)),          # interior
    (8, 10, (14, 12)),               # window leaves the frame at the top-left
    (470, 630, (468, 626)),          # bottom-right overhang
    (240, 3, (244, 9)),              # left edge
    (1, 320, (10, 318)),             # top edge, disk clipped
])
def test_config1_window_parity(gpu_pkg, oracle, cy, cx, guess):
    """480×640, dark disk tw=25 (r=12), default window 45 — BASELINE config 1 geometry."""
    f = disk_frame(480, 640, cy - 1, cx - 1, 12)
    check_step(gpu_pkg, oracle, f, 25, True, (45, 45), guess)


def test_generic_kernel_on_default_geometry(gpu_pkg, oracle):
    """The 45x45 / l=65 geometry normally dispatches to the window45 kernels; force the
    generic streaming kernel on the same frames so both stay parity-checked."""
    f = disk_frame(480, 640, 200, 300, 12)
    b = gpu_pkg.TrackerBatch(1, f.shape, 25, (45, 45), True)
    assert b.kernel_name == "dog_window45_argmax"
    b.set_option("window45", 0)
    assert b.kernel_name.startswith("dog_rect_argmax_generic")
    b.close()
    opts = {"window45": 0}
    check_step(gpu_pkg, oracle, f, 25, True, (45, 45), (198, 305), options=opts)
    check_step(gpu_pkg, oracle, disk_frame(480, 640, 5, 630, 12), 25, True, (45, 45), (10, 625), options=opts)


@pytest.mark.parametrize("tw,ws,darker", [
    (10, (21, 21), True), (10, (21, 33), False), (7, (15, 15), True), (33, (61, 61), True),
    (25, (44, 46), True), (25, (1, 1), True), (25, (91, 31), False), (40, (73, 73), True),
])
def test_kernel_lengths_and_window_shapes(gpu_pkg, oracle, tw, ws, darker):
    rng = np.random.default_rng(int(tw * 10) + ws[0])
    f = disk_frame(200, 260, 100, 130, int(tw) // 2, val=0 if darker else 255)
    f = np.clip(f.astype(int) + rng.integers(-6, 7, f.shape), 0, 255).astype(np.uint8)
    check_step(gpu_pkg, oracle, f, tw, darker, ws, (98, 134))


def test_noise_frames(gpu_pkg, oracle):
    rng = np.random.default_rng(11)
    for _ in range(4):
        f = rng.integers(0, 256, (150, 170)).astype(np.uint8)
        check_step(gpu_pkg, oracle, f, 10, True, (21, 21), (int(rng.integers(1, 151)), int(rng.integers(1, 171))))


def test_pitched_host_frame(gpu_pkg, oracle):
    big = np.full((130, 256), 128, np.uint8)
    big[:, :] = disk_frame(130, 256, 60, 100, 12)
    view = big[:, 3:203]                      # pitch 256, W 200, unaligned base
    assert not view.flags["C_CONTIGUOUS"]
    # Tracker() makes frames contiguous; exercise the pitched path at the batch level instead
    b = gpu_pkg.TrackerBatch(1, view.shape, 25, (45, 45), True)
    try:
        b.set_frames([view])
        fill = int(b.compute_fill()[0])
        assert fill == oracle.mode(view)
        ij, resp = b.step([[58, 101]])
        b.set_guess([[58, 101]])
        ij2, resp2 = b.track_host([[view]], mode="footprint")
        b.set_guess([[58, 101]])
        ij3, resp3 = b.track_host([[view]], mode="frames")
        ref = oracle.step(np.ascontiguousarray(view), fill, 25, True, (45, 45), (58, 101))
        assert tuple(ij[0]) == tuple(ij2[0, 0]) == tuple(ij3[0, 0]) == (ref.i, ref.j)
        assert resp[0] == resp2[0, 0] == resp3[0, 0]
    finally:
        b.close()


def test_float32_frames(gpu_pkg, oracle):
    """North-star layout: grayscale Float32 in [0,1]."""
    f8 = disk_frame(160, 200, 80, 100, 12)
    rng = np.random.default_rng(5)
    f8 = np.clip(f8.astype(int) + rng.integers(-9, 10, f8.shape), 0, 255).astype(np.uint8)
    f32 = (f8.astype(np.float32) / np.float32(255.0))
    fill = oracle.mode(f8)
    trk = gpu_pkg.Tracker(f32, 25, (45, 45), True)
    try:
        assert trk.fillvalue == fill
        ref = oracle.step(f8, fill, 25, True, (45, 45), (78, 104), want_map=True)
        assert trk((78, 104)) == (ref.i, ref.j)
        assert trk.step_resident((78, 104)) == (ref.i, ref.j)
        assert abs(trk.last_response - ref.resp) <= RTOL * ref.maxabs
        assert np.abs(trk.response_map((78, 104)) - ref.R).max() <= RTOL * ref.maxabs
    finally:
        trk.close()


def test_mode_kernel_tie_rule(gpu_pkg, oracle):
    rng = np.random.default_rng(2)
    cases = [np.array([[1, 2], [2, 1]], np.uint8), np.array([[3, 3, 9], [9, 9, 3]], np.uint8)]
    for _ in range(12):
        cases.append(rng.integers(0, 3, (int(rng.integers(1, 40)), int(rng.integers(1, 70)))).astype(np.uint8))
    cases.append(rng.integers(0, 256, (333, 517)).astype(np.uint8))
    half = np.zeros((64, 64), np.uint8); half[:, 32:] = 200          # exact 50/50 tie
    cases.append(half); cases.append(half[:, ::-1].copy())
    for f in cases:
        b = gpu_pkg.TrackerBatch(1, f.shape, 10, (21, 21), True)
        try:
            b.set_frames([f])
            assert int(b.compute_fill()[0]) == oracle.mode(f), f.shape
        finally:
            b.close()


@pytest.mark.parametrize("slow", [False, True])
def test_mode_fast_and_slow_paths(gpu_pkg, oracle, slow):
    """mode(frame): the counting pass (vector loads, lane-private histogram columns) decides alone unless two
    values tie for the maximum count; then — and with option mode_slow always — the last-position pass applies
    StatsBase's rule.  Frame widths that are not multiples of 16, u8 and f32 pixels, several frames per batch."""
    rng = np.random.default_rng(8)
    for (H, W) in [(37, 53), (64, 64), (120, 200), (9, 1000)]:
        frames = [rng.integers(0, 256, (H, W)).astype(np.uint8),                       # noise: near-ties likely
                  rng.integers(100, 104, (H, W)).astype(np.uint8),                     # four values
                  np.full((H, W), 7, np.uint8)]                                        # flat
        tie = np.zeros((H, W), np.uint8); tie[:, : W // 2] = 9; tie[:, W // 2: 2 * (W // 2)] = 200
        frames.append(tie)                                                             # exact tie (plus zeros if W odd)
        b = gpu_pkg.TrackerBatch(len(frames), (H, W), 10, (21, 21), True)
        try:
            b.set_option("mode_slow", int(slow))
            b.set_frames(frames)
            got = b.compute_fill()
            assert [int(x) for x in got] == [oracle.mode(f) for f in frames], (H, W)
            assert [int(x) for x in b.compute_fill()] == [int(x) for x in got]         # scratch left clean
        finally:
            b.close()
        f32 = [(f.astype(np.float32) / np.float32(255.0)) for f in frames]
        b = gpu_pkg.TrackerBatch(len(frames), (H, W), 10, (21, 21), True, dtype=np.float32)
        try:
            b.set_option("mode_slow", int(slow))
            b.set_frames(f32)
            assert [int(x) for x in b.compute_fill()] == [oracle.mode(f) for f in frames], (H, W)
        finally:
            b.close()


def test_blank_window_is_a_documented_near_tie(gpu_pkg, oracle):
    """A window that sees only the fill value has a flat response: the oracle's
    maximum is decided by 1e-17-level rounding noise, the GPU's (exactly zero)
    by findmax's first-element rule.  Documented, not compared."""
    f = np.full((100, 100), 128, np.uint8)
    ref = oracle.step(f, 128, 10, True, (21, 21), (50, 50))
    assert ref.near_tie(RTOL)
    trk = gpu_pkg.Tracker(f, 10, (21, 21), True)
    try:
        assert trk((50, 50)) == (40, 40)          # first element in column-major order
        assert trk.last_response == 0.0
    finally:
        trk.close()


def test_autodetect_window_480(gpu_pkg, oracle):
    """start_location = missing: window size .÷ 4 centred on the frame (:99-104)."""
    f = disk_frame(480, 640, 250, 300, 12)
    ws2 = (480 // 4, 640 // 4)
    ref = check_step(gpu_pkg, oracle, f, 25, True, ws2, (240, 320))
    assert (ref.i, ref.j) == (251, 301)


def test_autodetect_window_1080p(gpu_pkg, oracle):
    """BASELINE config 2 start-up: 271×481 outputs at 1080p, dense oracle (551 M MAC)."""
    f = disk_frame(1080, 1920, 539, 959, 12)
    ref = check_step(gpu_pkg, oracle, f, 25, True, (1080 // 4, 1920 // 4), (540, 960), check_map=False)
    assert (ref.i, ref.j) == (540, 960)
    g = disk_frame(1080, 1920, 450, 1100, 12)
    ref = check_step(gpu_pkg, oracle, g, 25, True, (270, 480), (540, 960), check_map=False)
    assert (ref.i, ref.j) == (451, 1101)


def test_config4_wide_halo_4k(gpu_pkg, oracle):
    """3840×2160, light target, tw=100 (l=245, halo 122): window 401 against the
    separable Float64 oracle (validated against the dense one in test_oracle.py),
    default window 173 against the dense reference-order oracle (1.8 G MAC)."""
    f = disk_frame(2160, 3840, 1000, 2000, 50, val=255)
    ref = check_step(gpu_pkg, oracle, f, 100, False, (401, 401), (1040, 1950), dense=False)
    assert (ref.i, ref.j) == (1001, 2001)
    ref = check_step(gpu_pkg, oracle, f, 100, False, (173, 173), (1010, 1990), dense=True)
    assert (ref.i, ref.j) == (1001, 2001)
    # window hanging over the frame corner with the wide halo
    g = disk_frame(2160, 3840, 30, 40, 50, val=255)
    check_step(gpu_pkg, oracle, g, 100, False, (173, 173), (60, 70), dense=False)


def test_full_frame_rect_1080p(gpu_pkg, oracle):
    """The full-frame DoG benchmark shape (1080×1920 outputs) vs the separable oracle."""
    f = disk_frame(1080, 1920, 700, 1234, 12)
    rng = np.random.default_rng(9)
    f = np.clip(f.astype(int) + rng.integers(-3, 4, f.shape), 0, 255).astype(np.uint8)
    b = gpu_pkg.TrackerBatch(1, f.shape, 25, (45, 45), True)
    try:
        b.set_frames([f])
        fill = int(b.compute_fill()[0])
        assert fill == oracle.mode(f)
        (oi, oj), raw, resp = b.rect_argmax(0, 0, 0, 1080, 1920)
        ref = oracle.rect(f, fill, 25, True, 0, 0, 1080, 1920, dense=False)
        assert (oi, oj) == (ref.i, ref.j) == raw
        assert abs(resp - ref.resp) <= RTOL * ref.maxabs
    finally:
        b.close()


@pytest.mark.parametrize("chunks", ["1", "2", "3", "5", "99"])
def test_marching_rect_kernel_chunkings(gpu_pkg, oracle, chunks):
    """dog_rect45_march: a strip is cut into `chunks` runs of 45-row batches; inside a run the row-pass
    intermediate is carried from batch to batch.  Every chunking must give the same response map (bit for
    bit — the same operations in the same order) and that map must match the oracle; the window hangs over
    two frame edges so the marching staging path meets rows and columns outside the frame."""
    rng = np.random.default_rng(21)
    f = rng.integers(0, 256, (300, 280)).astype(np.uint8)
    ws, guess = (231, 200), (190, 60)           # 231x201 outputs = 6x5 tiles; rows 75..305, cols -40..160
    trk = gpu_pkg.Tracker(f, 25, ws, False)
    try:
        trk.set_option("r45_chunks", int(chunks))
        fill = oracle.mode(f)
        assert trk.fillvalue == fill
        ref = oracle.step(f, fill, 25, False, ws, guess, dense=False, want_map=True)
        got = trk.step_resident(guess)
        resp = trk.last_response
        rmap = trk.response_map(guess)
        assert rmap.shape == ref.R.shape == (231, 201)
        assert np.abs(rmap.astype(np.float64) - ref.R).max() <= RTOL * ref.maxabs
        assert abs(resp - ref.resp) <= RTOL * ref.maxabs
        assert not ref.near_tie(RTOL)
        assert got == (ref.i, ref.j)
        # the published maximum is the maximum of the published map, at findmax's first position
        jj, ii = np.unravel_index(np.argmax(rmap.T), rmap.T.shape)
        assert resp == rmap[ii, jj]
        trk.set_option("r45_chunks", 99)                        # independent tiles
        assert np.array_equal(trk.response_map(guess), rmap)
    finally:
        trk.close()


def test_marching_rect_batch_of_frames(gpu_pkg, oracle):
    """Several 600x500 frames auto-detected in one launch (items = window x chunk x strip)."""
    rng = np.random.default_rng(4)
    H, W, n = 600, 500, 5
    cents = [(int(rng.integers(230, 370)), int(rng.integers(190, 310))) for _ in range(n)]
    frames = [disk_frame(H, W, cy, cx, 12) for cy, cx in cents]
    b = gpu_pkg.TrackerBatch(n, (H, W), 25, (H // 4, W // 4), True)
    try:
        b.set_frames(frames)
        b.compute_fill()
        ij, resp = b.step([[H // 2, W // 2]] * n)
        for v in range(n):
            ref = oracle.step(frames[v], 128, 25, True, (H // 4, W // 4), (H // 2, W // 2), dense=False)
            assert tuple(ij[v]) == (ref.i, ref.j) == (cents[v][0] + 1, cents[v][1] + 1)
            assert abs(resp[v] - ref.resp) <= RTOL * ref.maxabs
    finally:
        b.close()


# ---------------------------------------------------------------------------
def oracle_track(oracle, frames, tw, darker, ws, start_guess, autodetect=False):
    """The intended frame loop (src/PawsomeTracker.jl:159-169) driven by the oracle."""
    H, W = frames[0].shape
    fill = oracle.mode(frames[0])
    near = 0
    if autodetect:
        r = oracle.step(frames[0], fill, tw, darker, (H // 4, W // 4), (H // 2, W // 2), dense=False)
    else:
        r = oracle.step(frames[0], fill, tw, darker, ws, start_guess, dense=True)
    near += r.near_tie(RTOL)
    out = [(r.i, r.j)]
    for f in frames[1:]:
        r = oracle.step(f, fill, tw, darker, ws, out[-1], dense=True)
        near += r.near_tie(RTOL)
        out.append((r.i, r.j))
    return np.array(out), near


def test_config1_track_300_frames(gpu_pkg, oracle, synth):
    """BASELINE config 1: 480×640, 300 frames, dark disk tw=25, start given, default window."""
    start = (240, 320)
    r = 0.8 * min(start[0], start[1], 480 - start[0], 640 - start[1])
    tra = synth.spiral(r, 300, start, seed=0)
    vid = synth.SyntheticVideo(480, 640, tra, 25, True, fps=24.0)
    ts, ij = gpu_pkg.track(vid, start=0, stop=300 / 24.0, target_width=25,
                           start_location=gpu_pkg.CartesianIndex(*start), darker_target=True, fps=24)
    assert len(ts) == len(ij) == 300
    frames = [vid.frame(k) for k in range(300)]
    ref, near = oracle_track(oracle, frames, 25, True, (45, 45), start)
    assert near == 0
    np.testing.assert_array_equal(ij, ref)                      # bit-identical positions
    rmse = np.sqrt(np.mean(np.sum((ij - tra) ** 2, axis=1)))
    assert rmse < 1.0, rmse                                     # the reference's behavioural bar (README.md:24)
    assert ts[0] == 0 and abs(ts[-1] - 300 / 24.0) < 1e-12


def test_batch_equals_singles_and_all_paths_agree(gpu_pkg, oracle, synth):
    n, T, H, W = 6, 20, 240, 320
    vids = [synth.make_video(H=H, W=W, target_width=25, start_ij=(120, 160), seconds=10.0, fps=24.0, seed=s)
            for s in range(n)]
    steps = [[v.frame(t) for v in vids] for t in range(T)]
    with gpu_pkg.TrackerBatch(n, (H, W), 25, (45, 45), True) as b:
        b.set_frames(steps[0])
        fills = b.compute_fill()
        start = np.tile([120, 160], (n, 1))
        b.set_guess(start)
        ij_fp, r_fp = b.track_host(steps, mode="footprint")
        b.set_guess(start)
        ij_fr, r_fr = b.track_host(steps, mode="frames")
        # resident: all T×n frames in one device buffer (via torch), chained on the device
        import torch
        stack = torch.from_numpy(np.stack([np.stack(s) for s in steps])).cuda()      # (T, n, H, W) u8
        b.set_guess(start)
        ij_dev, r_dev = b.track_device(stack.data_ptr(), n * H * W, H * W, W, T)
        # step-by-step with host round trips
        b.set_guess(start)
        ij_st = []
        for t in range(T):
            b.set_frames(steps[t])
            o, _ = b.step(None)
            ij_st.append(o.copy())
    np.testing.assert_array_equal(ij_fp, ij_fr)
    np.testing.assert_array_equal(ij_fp, ij_dev)
    np.testing.assert_array_equal(ij_fp, np.stack(ij_st))
    np.testing.assert_array_equal(r_fp, r_fr)
    np.testing.assert_array_equal(r_fp, r_dev)
    for v in range(n):
        ref, near = oracle_track(oracle, [s[v] for s in steps], 25, True, (45, 45), (120, 160))
        assert near == 0 and fills[v] == 128
        np.testing.assert_array_equal(ij_fp[:, v], ref)


def test_zero_copy_pinned_host_frames(gpu_pkg, oracle, synth):
    """Pinned host frames: the chained kernel reads footprints over PCIe (zero-copy).  Must equal
    the staged footprint path (pageable frames) and the oracle loop, incl. windows leaving the frame."""
    import torch
    n, T, H, W = 5, 12, 200, 256
    vids = [synth.make_video(H=H, W=W, target_width=25, start_ij=(100, 128), seconds=10.0, fps=24.0, seed=10 + s)
            for s in range(n)]
    steps = [[v.frame(t) for v in vids] for t in range(T)]
    steps[3][0][:] = np.roll(steps[3][0], (-70, -100), axis=(0, 1))      # throw video 0 towards a corner
    pinned = torch.empty((T, n, H, W), dtype=torch.uint8, pin_memory=True)
    pnp = pinned.numpy()
    for t in range(T):
        for v in range(n):
            pnp[t, v] = steps[t][v]
    start = np.tile([100, 128], (n, 1))
    with gpu_pkg.TrackerBatch(n, (H, W), 25, (45, 45), True) as b:
        b.set_frames(steps[0]); b.compute_fill()
        b.set_guess(start)
        ij_pageable, r_pageable = b.track_host(steps, mode="footprint")
        b.set_guess(start)
        lc = b.launch_count
        ij_zc, r_zc = b.track_host([[pnp[t, v] for v in range(n)] for t in range(T)], mode="footprint")
        assert b.launch_count - lc == 1, "pinned frames must take the single-launch zero-copy path"
        # 5 videos at regular strides in one page-locked buffer: handed over as base + strides (region prefetch one step
        # ahead), lone-window kernel with 2 CTAs per window
        assert b.last_kernel == "dog_window45_cluster<2>"
        nxt, _ = b.step(None)                                          # chain state was left on the device
        # the same frames in an order that is NOT a regular stride pattern: frame-pointer table, 4 CTAs per window
        perm = [3, 0, 4, 1, 2]                                         # (every video has the same fill value, 128)
        b.set_guess(start)
        ij_tab, r_tab = b.track_host([[pnp[t, perm[v]] for v in range(n)] for t in range(T)], mode="footprint")
        assert b.last_kernel == "dog_window45_cluster<4>"
        np.testing.assert_array_equal(ij_tab, ij_zc[:, perm])
        np.testing.assert_array_equal(r_tab, r_zc[:, perm])
        b.set_option("cluster", 1)                                     # the same through the per-SM kernel
        b.set_guess(start)
        ij_ll, r_ll = b.track_host([[pnp[t, v] for v in range(n)] for t in range(T)], mode="footprint")
        assert b.last_kernel == "dog_window45_argmax"
        np.testing.assert_array_equal(ij_ll, ij_zc)
        np.testing.assert_array_equal(r_ll, r_zc)
        b.set_option("zero_copy", 0)
        b.set_guess(start)
        ij_staged, _ = b.track_host([[pnp[t, v] for v in range(n)] for t in range(T)], mode="footprint")
    np.testing.assert_array_equal(ij_zc, ij_pageable)
    np.testing.assert_array_equal(ij_zc, ij_staged)
    np.testing.assert_array_equal(r_zc, r_pageable)
    for v in range(n):
        ref, _ = oracle_track(oracle, [s[v] for s in steps], 25, True, (45, 45), (100, 128))
        np.testing.assert_array_equal(ij_zc[:, v], ref)


def test_track_autodetect_and_batch_api(gpu_pkg, oracle, synth):
    """start_location = missing → auto-detect then track; track_batch == per-video track."""
    vids = [synth.make_video(H=240, W=320, target_width=25, start_ij=(120, 160), seconds=10.0, fps=24.0, seed=s)
            for s in (3, 4)]
    singles = [gpu_pkg.track(v, stop=0.5, target_width=25, start_location=None, fps=24)[1] for v in vids]
    ts, ij = gpu_pkg.track_batch(vids, stop=0.5, target_width=25, start_location=None, fps=24)
    assert ij.shape == (12, 2, 2)
    for k, v in enumerate(vids):
        np.testing.assert_array_equal(ij[:, k], singles[k])
        ref, near = oracle_track(oracle, [v.frame(t) for t in range(12)], 25, True, (45, 45), None, autodetect=True)
        assert near == 0
        np.testing.assert_array_equal(singles[k], ref)


def test_config2_1080p_autodetect_then_track(gpu_pkg, oracle, synth):
    """BASELINE config 2 (shortened): one 1080p video, start_location = missing → auto-detect over the
    271×481 window centred on the frame (dog_rect45_march), then windowed tracking; positions identical to the
    oracle-driven loop and within 1 px RMS of the ground truth."""
    H, W, nfr = 1080, 1920, 60
    start = (540, 960)
    tra = synth.spiral(0.8 * 540, 3000, start, seed=0)[:nfr]          # the 3000-frame trajectory, first 60 frames
    vid = synth.SyntheticVideo(H, W, tra, 25, True, fps=24.0)
    ts, ij = gpu_pkg.track(vid, stop=nfr / 24.0, target_width=25, start_location=None, fps=24)
    assert len(ij) == nfr
    frames = [vid.frame(k) for k in range(nfr)]
    ref, near = oracle_track(oracle, frames, 25, True, (45, 45), None, autodetect=True)
    assert near == 0
    np.testing.assert_array_equal(ij, ref)
    assert np.sqrt(np.mean(np.sum((ij - tra) ** 2, axis=1))) < 1.0


def test_config5_segments_sar_start_fps(gpu_pkg, oracle, synth):
    """Segmented multi-file video, SAR=2, (x,y) start, non-zero start, fps resampling
    (test/test-basic-test.jl:43-49, 73-79, 91-104, 116-121; src/PawsomeTracker.jl:181-214)."""
    H, Wd, sar, src_fps, fps = 270, 960, 2, 24.0, 12.0
    start_disp = (135, 480)                                          # displayed (row, col)
    r = 0.8 * min(start_disp[0], start_disp[1], H - start_disp[0], Wd - start_disp[1])
    _, tra = synth.build_trajectory(r, src_fps, start_disp, seconds=12.0, seed=0)    # 289 source frames
    parts = synth.my_partition(len(tra), 3)
    segs = [synth.SyntheticVideo(H, Wd, tra[a:b + 1], 25, True, fps=src_fps, sar=sar) for a, b in parts]
    seg_start = [0.25, 0.0, 0.0]
    seg_stop = [(b - a + 1) / src_fps for a, b in parts]
    x0 = int(tra[parts[0][0] + 6, 1]); y0 = int(tra[parts[0][0] + 6, 0])      # where the target is at t=0.25 s
    ts, ij = gpu_pkg.track(segs, start=seg_start, stop=seg_stop, target_width=25,
                           start_location=[(x0, y0), None, None], darker_target=True, fps=fps)
    # oracle-driven replica of the same host logic
    exp = []
    end = None
    for s, (a, b), st, sp in zip(segs, parts, seg_start, seg_stop):
        n = int(round(fps * (sp - st)))
        idx = [int(np.floor((st + k / fps) * src_fps + 0.5)) for k in range(n)]
        idx = [i for i in idx if i < len(s)]
        frames = [s.frame(i) for i in idx]
        guess = (y0, int(round(x0 / sar))) if end is None else end
        out, near = oracle_track(oracle, frames, 25, True, (45, 45), guess)
        assert near == 0
        exp.append(out)
        end = tuple(out[-1])
    exp = np.concatenate(exp)
    np.testing.assert_array_equal(ij, exp)
    assert len(ts) == len(ij)
    step = (seg_stop[0] - seg_start[0]) / (int(round(fps * (seg_stop[0] - seg_start[0]))) - 1)
    np.testing.assert_allclose(np.diff(ts), step)                    # range(first, step=…, length=n) (:210)
    # accuracy in displayed coordinates: tracked columns × SAR (test/test-basic-test.jl:101-104,132)
    truth = []
    for s, st, sp in zip(segs, seg_start, seg_stop):
        n = int(round(fps * (sp - st)))
        truth += [s.traj[i] for i in [int(np.floor((st + k / fps) * src_fps + 0.5)) for k in range(n)] if i < len(s)]
    truth = np.array(truth)
    scaled = np.stack([ij[:, 0], ij[:, 1] * sar], axis=1)
    rmse = np.sqrt(np.mean(np.sum((scaled - truth) ** 2, axis=1)))
    assert rmse < 1.5, rmse        # column quantisation by SAR=2 adds up to 1 px


def test_segment_parallel_equals_serial(gpu_pkg, synth):
    """segments that bring their own start_location start independent chains which advance
    concurrently in one batch; the result must equal the reference-order serial loop (:202-206), including
    per-segment fills, the first-frame refinement, an auto-detected first segment and unequal segment lengths."""
    H, W, fps = 240, 320, 24.0
    _, tra = synth.build_trajectory(0.8 * 120, fps, (120, 160), seconds=6.0, seed=5)
    parts = synth.my_partition(len(tra), 5)
    segs = [synth.SyntheticVideo(H, W, tra[a:b + 1], 25, True, fps=fps) for a, b in parts]
    fr3 = np.stack([segs[3].frame(k) for k in range(len(segs[3]))])
    fr3[fr3 == 128] = 140                                                  # another background: the fill differs per segment
    segs[3] = gpu_pkg.ArrayVideo(fr3, fps=fps)
    stops = [(b - a + 1) / fps for a, b in parts]
    stops[1] = stops[1] * 0.5                                              # a shorter second segment
    for first_loc in (gpu_pkg.CartesianIndex(120, 160), None):
        locs = [first_loc, None, gpu_pkg.CartesianIndex(int(tra[parts[2][0], 0]), int(tra[parts[2][0], 1])), None,
                (int(tra[parts[4][0], 1]), int(tra[parts[4][0], 0]))]      # three chains: [0,1] [2,3] [4]
        kw = dict(start=[0.0] * 5, stop=stops, target_width=25, start_location=locs, darker_target=True, fps=fps)
        ts_s, ij_s = gpu_pkg.track(segs, **kw)
        ts_p, ij_p = gpu_pkg.track(segs, parallel=True, **kw)
        np.testing.assert_array_equal(ij_p, ij_s)
        np.testing.assert_array_equal(ts_p, ts_s)
        assert len(ij_s) > 100


def test_large_batch_1080p_properties(gpu_pkg):
    """BASELINE config 3 geometry at full size (256 × 1080p), size-independent checks:
    every video finds its disk centre; results do not depend on batch composition."""
    import torch
    n, H, W, T = 256, 1080, 1920, 3
    rng = np.random.default_rng(0)
    centres = np.stack([rng.integers(30, H - 30, (T, n)), rng.integers(30, W - 30, (T, n))], axis=-1)
    centres[1:] = centres[0] + rng.integers(-8, 9, (T - 1, n, 2))
    dev = torch.full((T, n, H, W), 128, dtype=torch.uint8, device="cuda")
    yy = torch.arange(-12, 13, device="cuda").view(-1, 1); xx = torch.arange(-12, 13, device="cuda").view(1, -1)
    mask = (yy * yy + xx * xx) <= 144
    for t in range(T):
        for v in range(n):
            cy, cx = int(centres[t, v, 0]) - 1, int(centres[t, v, 1]) - 1
            dev[t, v, cy - 12:cy + 13, cx - 12:cx + 13][mask] = 0
    torch.cuda.synchronize()          # torch renders on its own stream; the batch launches on another one
    with gpu_pkg.TrackerBatch(n, (H, W), 25, (45, 45), True) as b:
        b.bind_device_frames(dev.data_ptr(), H * W, W)
        assert (b.compute_fill() == 128).all()
        b.set_guess(centres[0] + rng.integers(-5, 6, (n, 2)))
        ij, resp = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
        np.testing.assert_array_equal(ij, centres)
        # permutation invariance: same videos in reverse order
        rev = dev.flip(1).contiguous()
        torch.cuda.synchronize()
        b.set_guess(ij[0][::-1].copy())
        ij2, resp2 = b.track_device(rev.data_ptr(), n * H * W, H * W, W, T)
        np.testing.assert_array_equal(ij2[:, ::-1], ij)
        np.testing.assert_array_equal(resp2[:, ::-1], resp)
    assert np.all(resp > 0.08)


@pytest.mark.parametrize("n,T", [(1, 1), (3, 5), (149, 4), (300, 3), (700, 2), (256, 9), (240, 7), (295, 5), (180, 6)])
def test_batch_sizes_exercise_cta_video_loop(gpu_pkg, oracle, n, T):
    """dog_window45_argmax hosts two videos per CTA and loops `v += 2·#CTAs`: cover one video, an odd
    count just above the SM count (lone second halves), and more videos than one wave holds; with
    1.18·#SMs <= n < 2·#SMs and T > 1 the chained call runs dog_window45_rot (windows hop between SMs while
    the empty slots rotate), which must agree with the per-step launches step by step.  Every
    video has its own frame content; spot-check a sample of videos against the oracle loop, all of them
    against ground truth, and the per-step path against the chained one."""
    import torch
    H, W = 128, 160
    rng = np.random.default_rng(n * 31 + T)
    c0 = np.stack([rng.integers(20, H - 20, n), rng.integers(20, W - 20, n)], axis=-1)
    cent = np.stack([c0 + rng.integers(-7, 8, (n, 2)) * (t > 0) for t in range(T)])      # (T, n, 2), 1-based
    cent = np.clip(np.cumsum(np.concatenate([c0[None], np.diff(cent, axis=0)]), axis=0), 1, [H, W])
    frames = np.full((T, n, H, W), 128, np.uint8)
    yy, xx = np.ogrid[0:H, 0:W]
    for t in range(T):
        for v in range(n):
            cy, cx = cent[t, v] - 1
            y0, y1, x0, x1 = max(0, cy - 12), min(H, cy + 13), max(0, cx - 12), min(W, cx + 13)
            sub = frames[t, v, y0:y1, x0:x1]
            sub[(yy[y0:y1] - cy) ** 2 + (xx[:, x0:x1] - cx) ** 2 <= 144] = 0
    dev = torch.from_numpy(frames).cuda()
    start = cent[0] + rng.integers(-4, 5, (n, 2))
    with gpu_pkg.TrackerBatch(n, (H, W), 25, (45, 45), True) as b:
        assert b.kernel_name == "dog_window45_argmax"
        b.bind_device_frames(dev.data_ptr(), H * W, W)
        b.set_fill(128)
        b.set_guess(start)
        ij, resp = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
        b.set_guess(start)
        per_step = []
        for t in range(T):
            b.bind_device_frames(dev.data_ptr() + t * n * H * W, H * W, W)
            o, _ = b.step(None)
            per_step.append(o.copy())
    np.testing.assert_array_equal(ij, np.stack(per_step))
    for v in sorted(set([0, n - 1, n // 2] + list(rng.integers(0, n, 4)))):
        ref, near = oracle_track(oracle, [frames[t, v] for t in range(T)], 25, True, (45, 45), tuple(start[v]))
        assert near == 0
        np.testing.assert_array_equal(ij[:, v], ref)
    # disks that are fully inside the frame are found exactly at their centre
    inside = (cent[..., 0] > 13) & (cent[..., 0] < H - 13) & (cent[..., 1] > 13) & (cent[..., 1] < W - 13)
    np.testing.assert_array_equal(ij[inside], cent[inside])


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_rotating_slots_equal_static_split(gpu_pkg, dtype):
    """dog_window45_rot vs dog_window45_argmax on the same 240-video, 11-step chain: identical positions AND
    bit-identical responses (same arithmetic, only the SM a window runs on changes); hand-off scratch left clean
    (a second chained call gives the same answer); the plain launch's co-residency handshake and its static fall-back."""
    import torch
    n, T, H, W = 240, 11, 96, 128
    rng = np.random.default_rng(77)
    base = rng.integers(0, 256, (T, n, H, W)).astype(np.uint8)          # noise: every response value is informative
    frames = base if dtype is np.uint8 else base.astype(np.float32) / np.float32(255.0)
    dev = torch.from_numpy(frames).cuda()
    start = np.stack([rng.integers(1, H + 1, n), rng.integers(1, W + 1, n)], axis=-1)
    with gpu_pkg.TrackerBatch(n, (H, W), 25, (45, 45), True, dtype=dtype) as b:
        b.bind_device_frames(dev.data_ptr(), H * W, W)
        b.compute_fill()
        b.set_guess(start)
        ij_rot, r_rot = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
        b.set_guess(start)
        ij_rot2, r_rot2 = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
        assert b.last_kernel == "dog_window45_rot"
        b.set_option("rot", 0)
        b.set_guess(start)
        ij_st, r_st = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
        assert b.last_kernel == "dog_window45_argmax"
        # the kernel's own fall-back: its co-residency handshake settles on the static schedule (what happens when the
        # device is shared and the CTAs cannot all be resident) — and the next launch rotates again
        b.set_option("rot", 3)
        b.set_guess(start)
        ij_fb, r_fb = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
        assert b.last_kernel == "dog_window45_rot"
        b.set_option("rot", 1)
        b.set_guess(start)
        ij_rot3, r_rot3 = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
        assert b.last_kernel == "dog_window45_rot"
    np.testing.assert_array_equal(ij_rot, ij_st)
    np.testing.assert_array_equal(r_rot, r_st)
    np.testing.assert_array_equal(ij_rot2, ij_st)
    np.testing.assert_array_equal(r_rot2, r_st)
    np.testing.assert_array_equal(ij_fb, ij_st)
    np.testing.assert_array_equal(r_fb, r_st)
    np.testing.assert_array_equal(ij_rot3, ij_st)
    np.testing.assert_array_equal(r_rot3, r_st)


def test_float32_frames_chained_and_batched(gpu_pkg, oracle, synth):
    """f32 frames through the chained (resident) path of the specialised kernel."""
    import torch
    n, T, H, W = 5, 6, 150, 170
    vids = [synth.make_video(H=H, W=W, target_width=25, start_ij=(75, 85), seconds=10.0, fps=24.0, seed=40 + s)
            for s in range(n)]
    f8 = np.stack([np.stack([v.frame(t) for v in vids]) for t in range(T)])
    f32 = f8.astype(np.float32) / np.float32(255.0)
    dev = torch.from_numpy(f32).cuda()
    with gpu_pkg.TrackerBatch(n, (H, W), 25, (45, 45), True, dtype=np.float32) as b:
        b.bind_device_frames(dev.data_ptr(), H * W, W)
        assert (b.compute_fill() == 128).all()
        b.set_guess(np.tile([75, 85], (n, 1)))
        ij, resp = b.track_device(dev.data_ptr(), n * H * W, H * W, W, T)
    for v in range(n):
        ref, near = oracle_track(oracle, [f8[t, v] for t in range(T)], 25, True, (45, 45), (75, 85))
        assert near == 0
        np.testing.assert_array_equal(ij[:, v], ref)


def test_guess_far_outside_frame(gpu_pkg, oracle):
    """A guess outside the frame is legal (the window then sees mostly the fill border); the result is clamped."""
    f = disk_frame(100, 120, 10, 110, 12)
    for guess in [(-30, 150), (1, 120), (140, -20), (30, 130)]:
        trk = gpu_pkg.Tracker(f, 25, (45, 45), True)
        try:
            ref = oracle.step(f, 128, 25, True, (45, 45), guess)
            got = trk(guess)
            assert 1 <= got[0] <= 100 and 1 <= got[1] <= 120
            if not ref.near_tie(RTOL):
                assert got == (ref.i, ref.j)
                assert abs(trk.last_response - ref.resp) <= RTOL * ref.maxabs
        finally:
            trk.close()


def test_host_decode_feeder_cv2(gpu_pkg, tmp_path, synth):
    """SURVEY §8(f) rank 1 ("next"): a real video file decoded on the host (OpenCV/FFmpeg → GRAY8, the role of
    `openvideo(…, AV_PIX_FMT_GRAY8)`, src/PawsomeTracker.jl:157) feeding track().  The codec is lossy, so the
    bar is the reference's behavioural one: RMSE < 1 px against the ground truth (README.md:24)."""
    cv2 = pytest.importorskip("cv2")
    H, W, nfr = 240, 320, 48
    tra = synth.spiral(0.8 * 120, 600, (120, 160), seed=3)[:nfr]
    vid = synth.SyntheticVideo(H, W, tra, 25, True, fps=24.0)
    path = str(tmp_path / "clip.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 24.0, (W, H), isColor=True)
    if not wr.isOpened():
        pytest.skip("OpenCV cannot write MJPG/AVI in this build")
    for k in range(nfr):
        wr.write(cv2.cvtColor(vid.frame(k), cv2.COLOR_GRAY2BGR))
    wr.release()
    src = gpu_pkg.CvVideo(path)
    assert len(src) == nfr and abs(src.fps - 24.0) < 1e-6
    ts, ij = gpu_pkg.track(path, stop=nfr / 24.0, target_width=25, start_location=gpu_pkg.CartesianIndex(120, 160), fps=24)
    assert len(ij) == nfr
    assert np.sqrt(np.mean(np.sum((ij - tra) ** 2, axis=1))) < 1.0
    # missing start → auto-detect on the decoded first frame
    ts2, ij2 = gpu_pkg.track(path, stop=0.5, target_width=25, start_location=None, fps=24)
    assert np.abs(ij2[0] - tra[0]).max() <= 1


def test_decode_feeder_pinned_ring_batch(gpu_pkg, tmp_path, synth):
    """several encoded files decoded concurrently by host threads into the ring of page-locked
    step-chunks (FrameFeeder) and tracked in one batch with zero-copy footprint reads; decoding is deterministic, so
    the batch must reproduce the per-file `track` exactly — including a file that is shorter than the others."""
    cv2 = pytest.importorskip("cv2")
    H, W, nfr = 240, 320, 40
    paths, tras = [], []
    for s_ in range(5):
        tra = synth.spiral(0.8 * 120, 600, (120, 160), seed=20 + s_)[:nfr - (7 if s_ == 3 else 0)]
        vid = synth.SyntheticVideo(H, W, tra, 25, True, fps=24.0)
        path = str(tmp_path / f"clip{s_}.avi")
        wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 24.0, (W, H), isColor=True)
        if not wr.isOpened():
            pytest.skip("OpenCV cannot write MJPG/AVI in this build")
        for k in range(len(tra)):
            wr.write(cv2.cvtColor(vid.frame(k), cv2.COLOR_GRAY2BGR))
        wr.release()
        paths.append(path); tras.append(tra)
    kw = dict(stop=nfr / 24.0, target_width=25, start_location=gpu_pkg.CartesianIndex(120, 160), fps=24)
    ts, ij = gpu_pkg.track_batch(paths, chunk_steps=6, decode_workers=3, **kw)
    assert ij.shape == (nfr - 7, 5, 2)                       # the batch stops with its shortest video (:162)
    for v, path in enumerate(paths):
        _, single = gpu_pkg.track(path, **kw)
        np.testing.assert_array_equal(ij[:, v], single[:nfr - 7])
        assert np.sqrt(np.mean(np.sum((ij[:, v] - tras[v][:nfr - 7]) ** 2, axis=1))) < 1.0
    # the pinned buffer type on its own
    pa = gpu_pkg.PinnedArray((3, 4, 5), np.uint8)
    pa.array[...] = 7
    assert pa.array.sum() == 7 * 60
    pa.close()


def test_diagnostic_video_and_device_downscale(gpu_pkg, tmp_path, synth):
    """`diagnostic_file` (src/diagnose.jl): one 360x640 frame per tracked frame with the
    frame downscaled on the device, dot and trail at the tracked point.  The downscale kernel is checked against
    the bilinear formula it implements (pixel-centre aligned; the reference's ImageTransformations is not vendored)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(12)
    H, W = 270, 480
    f = rng.integers(0, 256, (H, W)).astype(np.uint8)
    with gpu_pkg.TrackerBatch(2, (H, W), 25, (45, 45), True) as b:
        b.set_frames([f, 255 - f])
        got = b.downscale(90, 160)
    ys = (np.arange(90) + 0.5) * (H / 90) - 0.5
    xs = (np.arange(160) + 0.5) * (W / 160) - 0.5
    y0 = np.floor(ys).astype(int); x0 = np.floor(xs).astype(int)
    wy = (ys - y0)[:, None]; wx = (xs - x0)[None, :]
    yc0, yc1 = np.clip(y0, 0, H - 1), np.clip(y0 + 1, 0, H - 1)
    xc0, xc1 = np.clip(x0, 0, W - 1), np.clip(x0 + 1, 0, W - 1)
    for v, src in enumerate((f, 255 - f)):
        s_ = src.astype(np.float64)
        top = s_[yc0][:, xc0] * (1 - wx) + s_[yc0][:, xc1] * wx
        bot = s_[yc1][:, xc0] * (1 - wx) + s_[yc1][:, xc1] * wx
        ref = top * (1 - wy) + bot * wy
        assert np.abs(got[v].astype(np.float64) - ref).max() <= 0.5 + 1e-3
    # the diagnostics video of a short track
    nfr = 20
    tra = synth.spiral(0.8 * 135, 600, (135, 240), seed=1)[:nfr]
    vid = synth.SyntheticVideo(H, W, tra, 25, True, fps=24.0)
    out = str(tmp_path / "diag.avi")
    ts, ij = gpu_pkg.track(vid, stop=nfr / 24.0, target_width=25, start_location=gpu_pkg.CartesianIndex(135, 240),
                           fps=24, diagnostic_file=out)
    _, ij_plain = gpu_pkg.track(vid, stop=nfr / 24.0, target_width=25, start_location=gpu_pkg.CartesianIndex(135, 240), fps=24)
    np.testing.assert_array_equal(ij, ij_plain)                      # diagnostics do not disturb the track
    cap = cv2.VideoCapture(out)
    frames = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        frames.append(fr)
    assert len(frames) == nfr and frames[0].shape[:2] == (360, 640)
    last = cv2.cvtColor(frames[-1], cv2.COLOR_BGR2GRAY)
    pi, pj = int(round(ij[-1, 0] * 360 / H)) - 1, int(round(ij[-1, 1] * 640 / W)) - 1
    assert last[pi, pj] > 200                                        # white dot on the dark disk (MJPG is lossy)


def test_plain_c_consumer_of_the_abi(gpu_pkg, oracle, tmp_path):
    """A C program compiled against include/pawsome.h and linked to libpawsome_cuda.so (no Python, no torch):
    the drop-in boundary as a foreign binding sees it."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cabi_driver")
    libdir = os.path.dirname(gpu_pkg.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cabi_driver.c"),
                    "-o", exe, "-L", libdir, "-lpawsome_cuda", "-Wl,-rpath," + libdir], check=True, capture_output=True)
    T = 10
    r = subprocess.run([exe, str(T)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[-1] == "ok" and lines[1] == "fill 128"
    assert lines[0].split()[3] == "65" and lines[0].split()[5] == "45"
    frames = []
    for t in range(T):
        frames.append(disk_frame(240, 320, 100 + 3 * t - 1, 150 + 5 * t - 1, 12))
    ref, near = oracle_track(oracle, frames, 25, True, (45, 45), (97, 154))
    assert near == 0
    single = np.array([[int(x) for x in l.split()[2:4]] for l in lines if l.startswith("single")])
    batch = np.array([[int(x) for x in l.split()[2:6]] for l in lines if l.startswith("batch")])
    np.testing.assert_array_equal(single, ref)
    np.testing.assert_array_equal(batch[:, :2], ref)
    ref2, _ = oracle_track(oracle, frames, 25, True, (45, 45), (103, 146))
    np.testing.assert_array_equal(batch[:, 2:], ref2)
    np.testing.assert_array_equal(ref, np.array([[100 + 3 * t, 150 + 5 * t] for t in range(T)]))


def test_error_behaviour(gpu_pkg):
    f = np.full((64, 64), 128, np.uint8)
    b = gpu_pkg.TrackerBatch(2, (64, 64), 10, (21, 21), True)
    try:
        with pytest.raises(gpu_pkg.PawsomeError, match="PT_ERR_STATE"):
            b.step([[10, 10], [10, 10]])                          # no frame yet
        b.set_frames([f, f])
        with pytest.raises(gpu_pkg.PawsomeError, match="PT_ERR_STATE"):
            b.step([[10, 10], [10, 10]])                          # no fill yet
        b.set_fill(128)
        with pytest.raises(gpu_pkg.PawsomeError, match="PT_ERR_STATE"):
            b.step(None)                                          # no guess on the device
        with pytest.raises(ValueError, match="DimensionMismatch"):
            b.set_frames([f, np.zeros((32, 64), np.uint8)])
        with pytest.raises(gpu_pkg.PawsomeError, match="PT_ERR_ARG"):
            b.set_fill([300, 0])
        out, resp = b.step([[10, 10], [70, 70]])                  # guess outside the frame is legal (clamped result)
        assert out.min() >= 1 and out.max() <= 64
    finally:
        b.close()
    with pytest.raises(gpu_pkg.PawsomeError, match="PT_ERR_UNSUPPORTED"):
        gpu_pkg.TrackerBatch(1, (64, 64), 2000, (21, 21), True)   # kernel too long for shared memory
