"""CPU tests of the host-side mirror of the reference interface (no device):
window-size conventions, start-location forms, resampling, timestamps, the
synthetic-video recipe."""
from fractions import Fraction

import numpy as np
import pytest


def test_window_size_helpers(pkg):
    # fix_window_size: (w, h) → (h, w); l → (l, l)   (src/PawsomeTracker.jl:70-72)
    assert pkg.fix_window_size((30, 50)) == (50, 30)
    assert pkg.fix_window_size(45) == (45, 45)
    assert pkg.guess_window_size(25) == 45 and pkg.guess_window_size(100) == 173


class _V:
    def __init__(self, sar):
        self.sar = Fraction(sar)


def test_get_guess_three_forms(pkg):
    img = np.zeros((1080, 1920), np.uint8)
    assert pkg.get_guess(None, _V(1), img) == (540, 960)                          # missing → size .÷ 2 (:86-90)
    assert pkg.get_guess(pkg.CartesianIndex(7, 9), _V(2), img) == (7, 9)         # raw index, no SAR (:74-77)
    assert pkg.get_guess((101, 50), _V(2), img) == (50, 50)                      # 50.5 → 50 (ties to even) (:79-84)
    assert pkg.get_guess((103, 50), _V(2), img) == (50, 52)                      # 51.5 → 52
    assert pkg.get_guess((100, 40), _V(Fraction(4, 3)), img) == (40, 75)


def test_resampling_indices(pkg):
    from pawsometracker_jl_b200.api import _Resampled
    vid = pkg.ArrayVideo([np.full((4, 4), k, np.uint8) for k in range(50)], fps=10.0)
    r = _Resampled(vid, start=1.0, t=2.0, fps=5.0)       # source frames 10, 12, 14, … ; 10 output frames
    got = []
    while not r.eof():
        got.append(int(r.read()[0, 0]))
    assert got == list(range(10, 30, 2))
    r = _Resampled(vid, start=4.0, t=100.0, fps=10.0)    # runs into the end of the source
    n = 0
    while not r.eof():
        r.read(); n += 1
    assert n == 10


def test_partition_overlaps_by_one(pkg, synth):
    parts = synth.my_partition(241, 3)                      # test/test-basic-test.jl:43-49
    assert parts[0][0] == 0 and parts[-1][1] == 240
    for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
        assert a1 == b0
    assert synth.my_partition(10, 1) == [(0, 9)]


def test_spiral_recipe(pkg, synth):
    tra = synth.spiral(40.0, 241, (50, 50), seed=0)
    assert tra.shape == (241, 2) and tuple(tra[0]) == (50, 50)
    # uniform arc length: consecutive points are ~equidistant (jitter σ=1 per axis)
    d = np.linalg.norm(np.diff(tra, axis=0), axis=1)
    assert d.max() < 12 and np.abs(tra - 50).max() <= 40 + 6
    np.testing.assert_array_equal(tra, synth.spiral(40.0, 241, (50, 50), seed=0))   # seeded
    assert not np.array_equal(tra, synth.spiral(40.0, 241, (50, 50), seed=1))
    ts, tr = synth.build_trajectory(40.0, 24, (50, 50))
    assert len(ts) == 241 == len(tr) and ts[-1] == 10.0


def test_synthetic_video_frames(pkg, synth):
    v = synth.make_video(H=100, W=100, target_width=10, darker_target=True, start_ij=(50, 50))
    f = v.frame(0)
    assert f.dtype == np.uint8 and f.shape == (100, 100)
    assert f[49, 49] == 0 and f[0, 0] == 128 and (f == 0).sum() == 81           # filled disk of radius 5
    v2 = synth.make_video(H=100, W=200, target_width=10, darker_target=False, start_ij=(50, 100), sar=2)
    f2 = v2.frame(0)
    assert f2.shape == (100, 100) and f2[49, 49] == 255 and v2.stored_centre(0) == (50, 50)


def test_segment_length_mismatch_asserts(pkg, synth):
    v = synth.make_video()
    with pytest.raises(AssertionError, match="Array length mismatch"):
        pkg.track_segments([v, v], start=[0.0], stop=[1.0, 1.0], start_location=[None, None])


def test_diagnostics_scaling_and_trail(pkg, tmp_path):
    """src/diagnose.jl:26-32: ratio = size(buffer) ./ size(img); ij = round.(Int, point .* ratio); 100-point trail."""
    pytest.importorskip("cv2")
    d = pkg.Diagnose(str(tmp_path / "d.avi"), True)
    try:
        assert d.label == "d" and d.color == 255 and d.buffer.shape == (360, 640)
        d.update_ratio((1080, 1920))
        assert d.ratio == (1 / 3, 1 / 3)
        assert d.scaled((541, 961)) == (180, 320)
        assert pkg.Diagnose(str(tmp_path / "l.avi"), False).color == 0
    finally:
        d.close()


def test_segment_chains_split_at_explicit_start_locations(pkg):
    """`coalesce(loc, end_location)` (:204) links a segment to its predecessor only when its own
    start_location is missing."""
    files = list("abcdef")
    locs = [None, None, pkg.CartesianIndex(5, 6), None, (7, 8), pkg.CartesianIndex(1, 1)]
    chains = pkg.split_chains(files, [0.0] * 6, [1.0] * 6, locs)
    assert [[seg[0] for seg in c] for c in chains] == [[0, 1], [2, 3], [4], [5]]
    assert chains[0][0][4] is None and chains[1][0][4] == pkg.CartesianIndex(5, 6) and chains[1][1][4] is None
    assert [len(c) for c in pkg.split_chains(files, [0.0] * 6, [1.0] * 6, [None] * 6)] == [6]


def test_track_chunks_frame_order_for_every_kind_of_source(pkg, monkeypatch):
    """api._track_chunks (the chunked frame loop of track_one, src/PawsomeTracker.jl:163-169): sources that hand frames
    out by reference, sources that decode into the page-locked ring, and sources that switch between the two — every frame
    reaches the tracker exactly once and in order (host logic only: fake tracker, no device)."""
    import sys
    api = sys.modules[pkg.__name__ + ".api"]

    class FakePinned:
        def __init__(self, shape, dtype): self.array = np.zeros(shape, dtype)
        def close(self): pass

    monkeypatch.setattr(api, "PinnedArray", FakePinned)

    class FakeTrk:
        sz = (4, 6); img = np.zeros((4, 6), np.uint8)
        def track_frames(self, frames, guess):
            assert 1 <= len(frames) <= api.CHUNK_FRAMES_REF
            return np.array([[int(f[0, 0]), len(frames)] for f in frames], np.int32), None

    class Src:
        def __init__(self, n, by_ref): self.n, self.k, self.by_ref = n, 0, by_ref
        def eof(self): return self.k >= self.n
        def read_ref(self):
            if self.by_ref(self.k):
                f = np.full((4, 6), self.k % 251, np.uint8); self.k += 1
                return f
            return None
        def read(self, out=None):
            out[...] = self.k % 251; self.k += 1
            return out

    for by_ref in (lambda k: True, lambda k: False, lambda k: k < 100, lambda k: k % 2 == 0, lambda k: k > 70):
        for n in (1, 5, 64, 65, 300, 700):
            ind = [(0, 0)]
            api._track_chunks(FakeTrk(), Src(n, by_ref), n + 1, ind)
            assert [a for a, _ in ind[1:]] == [k % 251 for k in range(n)]
            ind = [(0, 0)]
            api._track_chunks(FakeTrk(), Src(n, by_ref), min(n, 40) + 1, ind)      # stop before the source ends
            assert [a for a, _ in ind[1:]] == [k % 251 for k in range(min(n, 40))]


def test_resampler_block_addresses_equal_per_frame_reads(pkg):
    """_Resampled.read_ref_block (vectorised frame selection + addresses for sources that hold all frames in one array)
    picks exactly the frames the per-frame path picks (`ffmpeg -ss start -t t -vf fps=fps`, src/PawsomeTracker.jl:155)."""
    import sys
    api = sys.modules[pkg.__name__ + ".api"]
    rng = np.random.default_rng(5)
    arr = rng.integers(0, 256, (97, 6, 8)).astype(np.uint8)
    for start, t, fps, src_fps in [(0.0, 4.0, 24.0, 24.0), (0.25, 3.0, 12.0, 24.0), (1.0, 10.0, 30.0, 24.0), (0.1, 2.5, 7.5, 25.0)]:
        a = api._Resampled(api.ArrayVideo(arr, fps=src_fps), start, t, fps)
        b = api._Resampled(api.ArrayVideo(arr, fps=src_fps), start, t, fps)
        per_frame = []
        while not a.eof():
            per_frame.append(a.read_ref())
        got = []
        while True:
            blk = b.read_ref_block(5)
            if blk is None:
                break
            addrs, shape, dt, pitch = blk
            assert tuple(shape) == (6, 8) and dt == np.uint8 and pitch == 8
            got.extend(int(x) for x in addrs)
        assert b.eof()
        assert got == [f.ctypes.data for f in per_frame]
    # a list of frames cannot be addressed as a block
    c = api._Resampled(api.ArrayVideo([arr[0], arr[1]], fps=24.0), 0.0, 1.0, 24.0)
    assert c.read_ref_block(4) is None and not c.eof()
