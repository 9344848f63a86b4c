"""Host-side mirror of the reference's public surface: `track(file; …)`,
`track(files; …)`, and the start-up helpers (src/PawsomeTracker.jl:64-214),
with the Tracker's arithmetic routed to libpawsome_cuda.so.

What is kept verbatim: keyword names and defaults, the three `start_location`
forms, window-size conventions, the timestamp construction, segment chaining.
What is different by necessity: there is no ffmpeg/VideoIO in this image, so a
"file" is any object implementing the small `VideoSource` protocol below
(`SyntheticVideo`, `ArrayVideo`, or `CvVideo` for real files through OpenCV).
`missing` is spelled `None`; `CartesianIndex(i, j)` is the class below.
"""
from __future__ import annotations

from fractions import Fraction
from typing import NamedTuple, Sequence

import math

import numpy as np

from .feeder import FrameFeeder, PinnedArray
from .tracker import Tracker, TrackerBatch, fix_window_size, guess_window_size

DEFAULT_MAX_DURATION_SECONDS = 86399.999  # src/PawsomeTracker.jl:19


class CartesianIndex(NamedTuple):
    """1-based (row, col) into the raw stored frame (src/PawsomeTracker.jl:118)."""
    i: int
    j: int


# ---------------------------------------------------------------------------
# video sources (the decode side stays on the host — BASELINE.json north_star)
# ---------------------------------------------------------------------------
class ArrayVideo:
    """Frames already in host memory: a (T, H, W) array or a list of HxW arrays."""

    def __init__(self, frames, fps: float = 24.0, sar=1):
        self.frames = frames
        self.fps = float(fps)
        self.sar = Fraction(sar)

    def __len__(self):
        return len(self.frames)

    def frame(self, k: int, out=None):
        f = self.frames[k]
        if out is None:
            return np.array(f, copy=True)
        np.copyto(out, f)
        return out

    def frame_ref(self, k: int):
        """The stored frame itself (no copy) when it can be handed to the tracker as it is, else None."""
        f = self.frames[k]
        return f if isinstance(f, np.ndarray) and f.ndim == 2 and f.strides[1] == f.itemsize else None

    def frame_addresses(self, idx: np.ndarray):
        """Addresses of the frames `idx` (an int64 array) as one int64 array + (shape, dtype, row pitch in elements), when
        the video is ONE (T, H, W) array with contiguous rows — no per-frame Python objects; else None."""
        a = self.frames
        if not (isinstance(a, np.ndarray) and a.ndim == 3 and a.strides[2] == a.itemsize and a.strides[1] % a.itemsize == 0):
            return None
        return a.ctypes.data + idx * a.strides[0], a.shape[1:], a.dtype, a.strides[1] // a.itemsize


class CvVideo:
    """A real video file decoded on the host with OpenCV's FFmpeg backend to
    GRAY8 (the role of `openvideo(…, target_format = AV_PIX_FMT_GRAY8)`,
    src/PawsomeTracker.jl:157)."""

    def __init__(self, path: str):
        import cv2  # host decode only
        self._cv2 = cv2
        self.cap = cv2.VideoCapture(path)
        if not self.cap.isOpened():
            raise FileNotFoundError(path)
        self.fps = float(self.cap.get(cv2.CAP_PROP_FPS)) or 24.0
        num = self.cap.get(cv2.CAP_PROP_SAR_NUM) or 1
        den = self.cap.get(cv2.CAP_PROP_SAR_DEN) or 1
        self.sar = Fraction(int(num), int(den)) if num > 0 and den > 0 else Fraction(1)
        self._n = int(self.cap.get(cv2.CAP_PROP_FRAME_COUNT))
        self._next = 0

    def __len__(self):
        return self._n

    def frame(self, k: int, out=None):
        if k != self._next:
            self.cap.set(self._cv2.CAP_PROP_POS_FRAMES, k)
        ok, bgr = self.cap.read()
        if not ok:
            raise EOFError
        self._next = k + 1
        g = self._cv2.cvtColor(bgr, self._cv2.COLOR_BGR2GRAY)
        if out is None:
            return g
        np.copyto(out, g)
        return out

    def frame_ref(self, k: int):
        """The freshly decoded GRAY8 frame: it is a new array anyway, the tracker can read it in place."""
        return self.frame(k)


def open_video(file):
    if isinstance(file, str):
        return CvVideo(file)
    if hasattr(file, "frame") and hasattr(file, "fps"):
        return file
    raise TypeError("file must be a path or a VideoSource (SyntheticVideo / ArrayVideo / CvVideo)")


def aspect_ratio(vid) -> Fraction:
    """VideoIO.aspect_ratio(vid) — used at src/PawsomeTracker.jl:80."""
    return Fraction(getattr(vid, "sar", 1))


class _Resampled:
    """`ffmpeg -ss start -i file -t t -vf fps=fps` (src/PawsomeTracker.jl:155):
    output frame k shows the source frame nearest to time start + k/fps; the
    stream ends after round(t·fps) frames or at the end of the source."""

    def __init__(self, vid, start: float, t: float, fps: float):
        self.vid, self.start, self.fps = vid, float(start), float(fps)
        self.limit = int(round(float(t) * float(fps)))
        self.k = 0

    def _src_index(self, k: int) -> int:
        return math.floor((self.start + k / self.fps) * self.vid.fps + 0.5)     # (math.floor: 10x cheaper than np.floor per frame)

    def eof(self) -> bool:
        return self.k >= self.limit or self._src_index(self.k) >= len(self.vid)

    def read(self, out=None):
        if self.eof():
            raise EOFError
        f = self.vid.frame(self._src_index(self.k), out)
        self.k += 1
        return f

    def read_ref_block(self, kmax: int):
        """The next ≤ kmax frames by ADDRESS in one vectorised step (sources that hold all frames in one array): returns
        (int64 address array, shape, dtype, pitch) and advances, or None when the source cannot do that."""
        get = getattr(self.vid, "frame_addresses", None)
        if get is None or self.k >= self.limit:
            return None
        k = np.arange(self.k, min(self.limit, self.k + int(kmax)), dtype=np.float64)
        idx = np.floor((self.start + k / self.fps) * self.vid.fps + 0.5).astype(np.int64)     # _src_index, vectorised
        idx = idx[idx < len(self.vid)]
        if idx.size == 0:
            return None
        r = get(idx)
        if r is None:
            return None
        self.k += int(idx.size)
        return r

    def read_ref(self):
        """Next frame by reference when the source can give one (no copy into the tracker's buffer), else None
        (the caller then falls back to `read(out=…)`, the `read!` of :166)."""
        get = getattr(self.vid, "frame_ref", None)
        if get is None:
            return None
        if self.eof():
            raise EOFError
        f = get(self._src_index(self.k))
        if f is not None:
            self.k += 1
        return f


# ---------------------------------------------------------------------------
# diagnostics video (src/diagnose.jl)
# ---------------------------------------------------------------------------
DIAGNOSTIC_VIDEO_SIZE = (360, 640)      # src/diagnose.jl:2
TRACE_BUFFER_SIZE = 100                 # src/diagnose.jl:3


class Diagnose:
    """`Diagnose(file, darker_target)` — src/diagnose.jl:5-24.  Per frame (:30-38): the tracked point scaled to
    the 360x640 buffer joins a 100-point trail, the frame is `imresize!`d into the buffer (on the device:
    pt_batch_downscale), label, dot (radius 2) and trail are drawn in white for dark targets / black for light
    ones, and the buffer goes to the encoder (host, OpenCV — the reference uses VideoIO's writer)."""

    def __init__(self, file: str, darker_target: bool, fps: float = 24.0):
        import os
        import cv2
        self._cv2 = cv2
        self.label = os.path.splitext(os.path.basename(file))[0]
        self.buffer = np.zeros(DIAGNOSTIC_VIDEO_SIZE, np.uint8)
        self.color = 255 if darker_target else 0
        ext = os.path.splitext(file)[1].lower()
        fourcc = cv2.VideoWriter_fourcc(*("MJPG" if ext == ".avi" else "mp4v"))
        self.writer = cv2.VideoWriter(file, fourcc, float(fps), (DIAGNOSTIC_VIDEO_SIZE[1], DIAGNOSTIC_VIDEO_SIZE[0]),
                                      isColor=False)
        if not self.writer.isOpened():
            raise OSError(f"cannot open {file} for writing")
        self.trace = []                     # CircularBuffer{CartesianIndex{2}}(100)
        self.ratio = (1.0, 1.0)
        self.frames_written = 0

    def update_ratio(self, sz):             # update_ratio! (:26-28)
        self.ratio = (DIAGNOSTIC_VIDEO_SIZE[0] / sz[0], DIAGNOSTIC_VIDEO_SIZE[1] / sz[1])

    def scaled(self, point):
        return (int(round(point[0] * self.ratio[0])), int(round(point[1] * self.ratio[1])))   # round.(Int, point .* ratio)

    def __call__(self, trckr, point):       # (dia::Diagnose)(img, point) (:30-38)
        cv2 = self._cv2
        ij = self.scaled(point)
        self.trace.append(ij)
        if len(self.trace) > TRACE_BUFFER_SIZE:
            self.trace.pop(0)
        self.buffer[...] = trckr.downscaled(*DIAGNOSTIC_VIDEO_SIZE)                           # imresize!(dia.buffer, img)
        cv2.putText(self.buffer, self.label, (20, 40), cv2.FONT_HERSHEY_SIMPLEX, 0.7, int(self.color), 1, cv2.LINE_AA)
        cv2.circle(self.buffer, (ij[1] - 1, ij[0] - 1), 2, int(self.color), -1)               # CirclePointRadius(ij, 2)
        if len(self.trace) > 1:
            pts = np.array([[j - 1, i - 1] for i, j in self.trace], np.int32).reshape(-1, 1, 2)
            cv2.polylines(self.buffer, [pts], False, int(self.color), 1)                      # Path(dia.trace)
        self.writer.write(self.buffer)
        self.frames_written += 1

    def close(self):
        self.writer.release()


class _Dont:                                # struct Dont (:42-46)
    def update_ratio(self, sz):
        pass

    def __call__(self, trckr, point):
        pass

    def close(self):
        pass


def diagnose(file, darker_target):
    return _Dont() if file is None else Diagnose(file, darker_target)


# ---------------------------------------------------------------------------
# start-up helpers
# ---------------------------------------------------------------------------
def get_guess(start_location, vid, img):
    """src/PawsomeTracker.jl:74-90"""
    if start_location is None:                                   # ::Missing  (:86-90)
        return (img.shape[0] // 2, img.shape[1] // 2)
    if isinstance(start_location, CartesianIndex):              # (:74-77)
        return (int(start_location.i), int(start_location.j))
    x, y = start_location                                       # (x, y) displayed px (:79-84)
    sar = aspect_ratio(vid)
    return (int(y), int(round(Fraction(int(x)) / sar)))          # round(Rational): ties to even, as Julia


def get_start_ij_and_tracker(start_location, vid, img, target_width, window_size, darker_target, device=0):
    """src/PawsomeTracker.jl:92-107"""
    guess = get_guess(start_location, vid, img)
    if start_location is None:
        sz = img.shape
        window_size2 = (sz[0] // 4, sz[1] // 4)                 # "this greatly affects processing time!" (:102)
        trckr = Tracker(img, target_width, window_size2, darker_target, device)   # auto-detection pass (:103)
        ij = trckr(guess)
        fill = trckr.fillvalue
        trckr.close()
        trckr = Tracker(img, target_width, window_size, darker_target, device, _fillvalue=fill)    # (:105; mode(img) again = fill)
        return trckr, ij
    trckr = Tracker(img, target_width, window_size, darker_target, device)
    ij = trckr(guess)                                           # the first frame is refined, not trusted (:95)
    return trckr, ij


CHUNK_FRAMES = 64        # frames per chained library call of track_one (the depth of the page-locked decode ring)
CHUNK_FRAMES_REF = 256   # … when the source hands its frames out by reference (no ring to allocate)
CHUNK_FRAMES_BLOCK = 1024  # … when it hands out whole blocks of frame addresses (frames in one array)


def _track_chunks(trckr, vid, n, indices):
    """The frame loop of track_one without a diagnostics writer: frames are taken CHUNK_FRAMES at a time and each
    chunk is ONE library call (Tracker.track_frames → pt_batch_track_host: the chain ij[k] = trckr(ij[k-1]) runs
    inside the library; page-locked frames are read in place by one chained kernel launch).  A source that hands out
    contiguous frames by reference (`frame_ref`: frames already in host memory) is tracked in place; any other
    source decodes into a ring of page-locked frames — the `read!(vid, trckr.img.data)` target of :166, K deep — so
    the kernels read the decoder's output without a staging copy."""
    ring = None
    H, W = trckr.sz
    dtype = trckr.img.dtype
    contiguous = (W * dtype.itemsize, dtype.itemsize)
    # sources that hold every frame in one array: whole chunks by address, no per-frame Python work
    while len(indices) < n and not vid.eof():
        blk = vid.read_ref_block(min(CHUNK_FRAMES_BLOCK, n - len(indices))) if hasattr(vid, "read_ref_block") else None
        if blk is None:
            break
        addrs, shape, dt, pitch = blk
        if tuple(shape) != (H, W) or dt != dtype:
            raise ValueError(f"DimensionMismatch: frames are {tuple(shape)} {dt}, tracker was built for {(H, W)} {dtype}")
        ij, _ = trckr.track_addresses(addrs, pitch, indices[-1])
        indices.extend(map(tuple, ij.tolist()))
    try:
        pending = None                       # a consumed frame that needs a ring slot of the NEXT chunk
        while (pending is not None or not vid.eof()) and len(indices) < n:
            frames = []
            nring = 0                        # ring slots used by this chunk
            while len(indices) + len(frames) < n and len(frames) < CHUNK_FRAMES_REF and nring < CHUNK_FRAMES:
                if pending is not None:
                    f, pending = pending, None
                elif vid.eof():
                    break
                else:
                    f = vid.read_ref()
                if f is None or f.shape != (H, W) or f.dtype != dtype or f.strides != contiguous:
                    if ring is None:
                        if frames:           # frames by reference so far: track them first, then start the ring
                            pending = f if f is not None else False
                            break
                        ring = PinnedArray((CHUNK_FRAMES, H, W), dtype)
                    slot = ring.array[nring]
                    nring += 1
                    if f is not None and f is not False:
                        np.copyto(slot, f)
                    else:
                        vid.read(out=slot)                      # read!(vid, trckr.img.data) (:166)
                    f = slot
                frames.append(f)
            if pending is False:
                pending = None               # (nothing was consumed: the next chunk reads it into the ring)
                if ring is None:
                    ring = PinnedArray((CHUNK_FRAMES, H, W), dtype)
            elif pending is not None and ring is None:
                ring = PinnedArray((CHUNK_FRAMES, H, W), dtype)
            if not frames:
                continue
            ij, _ = trckr.track_frames(frames, indices[-1])     # (:167) for the whole chunk
            indices.extend(map(tuple, ij.tolist()))
    finally:
        if ring is not None:
            ring.close()


def track_one(file, start, stop, target_width, start_location, window_size, darker_target, fps, dia=None, device=0):
    """src/PawsomeTracker.jl:148-174 with the intended loop body (:162, :167):
    `while !eof(vid) && last_frame < n`, `indices[k] = trckr(indices[k-1])`."""
    t = stop - start
    n = int(round(fps * t))
    if n < 1:
        raise ValueError(f"no frame to track: round(fps * (stop - start)) = {n} (the reference fails on `read` here)")
    ts = np.linspace(start, stop, n)                            # range(start, stop, n) (:152)
    dia = dia if dia is not None else _Dont()
    vid = _Resampled(open_video(file), start, t, fps)
    img = vid.read()
    dia.update_ratio(img.shape)                                 # (:160)
    trckr, ij = get_start_ij_and_tracker(start_location, vid.vid, img, target_width, window_size, darker_target, device)
    indices = [ij]
    try:
        dia(trckr, ij)
        own = trckr.img                                         # the tracker's private buffer (trckr.img.data)
        if isinstance(dia, _Dont):
            _track_chunks(trckr, vid, n, indices)               # no per-frame side channel: chained chunks
        while not vid.eof() and len(indices) < n:
            f = vid.read_ref()                                  # a frame the source already holds in host memory:
            if f is not None and f.shape == trckr.sz and f.dtype == own.dtype:
                trckr.img = f                                   # … is tracked in place (no 2 MB copy per frame),
            else:
                if f is not None:
                    np.copyto(own, f)
                else:
                    vid.read(out=own)                           # else read!(vid, trckr.img.data) (:166)
                trckr.img = own
            indices.append(trckr(indices[-1]))                  # (:167)
            dia(trckr, indices[-1])                             # (:168)
    finally:
        trckr.close()
    last = len(indices)
    return ts[:last], np.asarray(indices, np.int64).reshape(last, 2)


# ---------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------
def track(file, *, start=0, stop=DEFAULT_MAX_DURATION_SECONDS, target_width=25, start_location=None,
          window_size=None, darker_target=True, fps=24, diagnostic_file=None, device=0, parallel=False):
    """`track(file; …)` (src/PawsomeTracker.jl:130-146) or, when `file` is a list,
    the segmented `track(files; …)` (:181-214).  Returns (ts, ij) with ij an
    (n, 2) array of 1-based (row, col).  `parallel` (lists only): see track_segments."""
    if isinstance(file, (list, tuple)):
        return track_segments(file, start=start, stop=stop, target_width=target_width,
                              start_location=start_location, window_size=window_size,
                              darker_target=darker_target, fps=fps, diagnostic_file=diagnostic_file, device=device,
                              parallel=parallel)
    if window_size is None:
        window_size = guess_window_size(target_width)
    window_size = fix_window_size(window_size)
    dia = diagnose(diagnostic_file, darker_target)              # diagnose(...) do dia … end (:143-145)
    try:
        return track_one(file, start, stop, target_width, start_location, window_size, darker_target, fps, dia, device)
    finally:
        dia.close()


def track_segments(files: Sequence, *, start=None, stop=None, target_width=25, start_location=None,
                   window_size=None, darker_target=True, fps=24, diagnostic_file=None, device=0,
                   parallel: bool = False):
    """`track(files::AbstractVector; …)` — src/PawsomeTracker.jl:181-214.

    parallel=True (no equivalent in the reference): a segment that comes with its own
    `start_location` does not depend on the segment before it (`coalesce`, :204), so the file list splits into
    independent CHAINS of segments; the chains advance concurrently as the videos of one TrackerBatch.  The
    result is identical to the serial loop."""
    if parallel:
        if diagnostic_file is not None:
            raise ValueError("diagnostic_file needs the serial segment loop (one writer, frames in order)")
        return _track_segments_parallel(files, start, stop, target_width, start_location, window_size,
                                        darker_target, fps, device)
    start, stop, start_location = _segment_args(files, start, stop, start_location)
    if window_size is None:
        window_size = guess_window_size(target_width)
    window_size = fix_window_size(window_size)
    tss, ijs = [], []
    end_location = None
    dia = diagnose(diagnostic_file, darker_target)                                # one writer for all segments (:200)
    try:
        for f, t_start, t_stop, loc in zip(files, start, stop, start_location):
            loc = loc if loc is not None else end_location                        # coalesce (:204)
            ts_i, ij_i = track_one(f, t_start, t_stop, target_width, loc, window_size, darker_target, fps, dia, device)
            tss.append(ts_i)
            ijs.append(ij_i)
            end_location = CartesianIndex(int(ij_i[-1, 0]), int(ij_i[-1, 1]))     # (:206)
    finally:
        dia.close()
    n = sum(len(t) for t in tss)
    step = (tss[0][1] - tss[0][0]) if len(tss[0]) > 1 else 0.0
    ts = tss[0][0] + step * np.arange(n)                                          # range(first, step=…, length=n) (:210)
    return ts, np.concatenate(ijs, axis=0)


def _segment_args(files, start, stop, start_location):
    nfiles = len(files)
    start = [0.0] * nfiles if start is None or np.isscalar(start) and start == 0 else list(start)
    stop = ([DEFAULT_MAX_DURATION_SECONDS] * nfiles
            if stop is None or np.isscalar(stop) and stop == DEFAULT_MAX_DURATION_SECONDS else list(stop))
    start_location = [None] * nfiles if start_location is None else list(start_location)
    if not (nfiles == len(start) == len(stop) == len(start_location)):            # @assert (:193)
        raise AssertionError(f"Array length mismatch: files={nfiles}, start={len(start)}, stop={len(stop)}, "
                             f"start_location={len(start_location)}")
    return start, stop, start_location


class _Chain:
    """Consecutive segments linked by `coalesce(loc, end_location)` (:204): only the first has its own start."""

    def __init__(self, segs, fps):
        self.segs = segs                      # [(file index, file, t_start, t_stop, loc)]
        self.fps = fps
        self.k = -1                           # current segment
        self.vid = None
        self.n = 0
        self.count = 0
        self.out = {}                         # file index -> list of (i, j)
        self.last = None                      # previous result (chain state)
        self.done = False
        self.placeholder = None               # last frame seen: stands in once the chain is exhausted
        self._open_next()

    def _open_next(self):
        self.k += 1
        if self.k >= len(self.segs):
            self.done = True
            return
        idx, f, t0, t1, loc = self.segs[self.k]
        t = t1 - t0
        self.n = int(round(self.fps * t))
        self.vid = _Resampled(open_video(f), t0, t, self.fps)
        self.count = 0
        self.out[idx] = []

    def next_frame(self):
        """(frame, is_first_frame_of_segment) or None when the chain is exhausted."""
        while not self.done:
            first = self.count == 0
            if (first or self.count < self.n) and not self.vid.eof():      # read(vid) :159, loop condition :162
                f = self.vid.read_ref()                                    # by reference when the source allows
                return (f if f is not None else self.vid.read()), first
            if first:
                # a segment that yields no frame at all: the serial path fails in `read(vid)` (:159) — so does this one
                raise EOFError(f"segment {self.segs[self.k][0]} has no frame in [{self.segs[self.k][2]}, {self.segs[self.k][3]})")
            self._open_next()
        return None

    def record(self, ij):
        idx = self.segs[self.k][0]
        self.out[idx].append((int(ij[0]), int(ij[1])))
        self.last = (int(ij[0]), int(ij[1]))
        self.count += 1


def split_chains(files, start, stop, start_location):
    """Independent chains of a segmented video: a segment with its own start_location starts a new chain, a
    `missing` one continues from the end of its predecessor (`coalesce(loc, end_location)`, :204).
    Returns [[(file index, file, t_start, t_stop, loc), …], …]."""
    groups = []
    for i, (f, t0, t1, loc) in enumerate(zip(files, start, stop, start_location)):
        if i == 0 or loc is not None:
            groups.append([])
        groups[-1].append((i, f, t0, t1, loc))
    return groups


def _track_segments_parallel(files, start, stop, target_width, start_location, window_size, darker_target, fps, device):
    start, stop, start_location = _segment_args(files, start, stop, start_location)
    if window_size is None:
        window_size = guess_window_size(target_width)
    window_size = fix_window_size(window_size)
    chains = [_Chain(g, fps) for g in split_chains(files, start, stop, start_location)]
    nc = len(chains)
    cur = [c.next_frame() for c in chains]
    if any(x is None for x in cur):
        raise EOFError("a segment chain has no frame")
    shapes = {x[0].shape for x in cur}
    if len(shapes) != 1:
        raise ValueError("parallel segment tracking needs segments of one frame size")
    H, W = cur[0][0].shape
    batch = TrackerBatch(nc, (H, W), target_width, window_size, darker_target, dtype=cur[0][0].dtype, device=device)
    fills = np.zeros(nc, np.int32)
    try:
        while True:
            active = [x is not None for x in cur]
            if not any(active):
                break
            # finished chains keep their last frame as a placeholder; their results are ignored
            frames = []
            for v, x in enumerate(cur):
                if x is not None:
                    frames.append(x[0])
                else:
                    frames.append(chains[v].placeholder)
            new_seg = [x is not None and x[1] for x in cur]
            if any(new_seg):
                # a new Tracker per segment: fill = mode of the segment's first frame (:47, :94)
                batch.set_frames(frames)
                f_all = batch.compute_fill()
                for v in range(nc):
                    if new_seg[v]:
                        fills[v] = f_all[v]
                batch.set_fill(fills)
            guess = np.zeros((nc, 2), np.int32)
            auto = [False] * nc
            for v, c in enumerate(chains):
                if cur[v] is None:
                    guess[v] = c.last if c.last is not None else (1, 1)
                elif new_seg[v]:
                    loc = c.segs[c.k][4]
                    loc = loc if loc is not None else (CartesianIndex(*c.last) if c.last is not None else None)   # coalesce (:204)
                    guess[v] = get_guess(loc, c.vid.vid, cur[v][0])
                    auto[v] = loc is None
                else:
                    guess[v] = c.last
            if any(auto):
                # start_location = missing on the very first segment: auto-detect window size .÷ 4 (:99-105)
                batch.set_window((H // 4, W // 4))
                out_a, _ = batch.step(guess)
                batch.set_window(window_size)
            batch.set_guess(guess)
            out, _ = batch.track_host([frames], mode="footprint")
            for v, c in enumerate(chains):
                if cur[v] is None:
                    continue
                c.placeholder = frames[v]
                c.record(out_a[v] if auto[v] else out[0, v])
            cur = [c.next_frame() if x is not None else None for c, x in zip(chains, cur)]
    finally:
        batch.close()
    tss, ijs = [], []
    for i, (t0, t1) in enumerate(zip(start, stop)):
        n = int(round(fps * (t1 - t0)))
        ts_i = np.linspace(t0, t1, n)
        rec = next(c.out[i] for c in chains if i in c.out)
        tss.append(ts_i[:len(rec)])
        ijs.append(np.asarray(rec, np.int64).reshape(len(rec), 2))
    n = sum(len(t) for t in tss)
    step = (tss[0][1] - tss[0][0]) if len(tss[0]) > 1 else 0.0
    ts = tss[0][0] + step * np.arange(n)                                          # (:209-210)
    return ts, np.concatenate(ijs, axis=0)


def track_batch(files: Sequence, *, start=0, stop=DEFAULT_MAX_DURATION_SECONDS, target_width=25,
                start_location=None, window_size=None, darker_target=True, fps=24, device=0,
                chunk_steps: int | None = None, decode_workers: int | None = None):
    """Batched counterpart with no equivalent in the reference: `track` over
    many independent videos of identical geometry advanced in lock-step, one
    CTA group per (video, window) per launch.  Per-video results are identical
    to calling `track` on each video — up to the length of the SHORTEST video: the
    batch advances in lock-step and stops when any video reaches its end (`while
    !eof(vid) && last_frame < n`, :162, with eof = any video's eof); track longer
    videos separately (or in a second batch) if their tails matter.  start_location:
    None, one location, or one per video.  The videos are decoded concurrently by `decode_workers` host threads into a ring of
    page-locked step-chunks (feeder.FrameFeeder) that the kernels read without a staging copy."""
    nv = len(files)
    if window_size is None:
        window_size = guess_window_size(target_width)
    window_size = fix_window_size(window_size)
    locs = list(start_location) if isinstance(start_location, list) else [start_location] * nv
    t = stop - start
    n = int(round(fps * t))
    if n < 1:
        raise ValueError(f"no frame to track: round(fps * (stop - start)) = {n}")
    ts = np.linspace(start, stop, n)
    vids = [_Resampled(open_video(f), start, t, fps) for f in files]
    first = [v.read() for v in vids]
    H, W = first[0].shape
    guess = np.array([get_guess(loc, v.vid, img) for loc, v, img in zip(locs, vids, first)], np.int32)
    batch = TrackerBatch(nv, (H, W), target_width, window_size, darker_target, dtype=first[0].dtype, device=device)
    try:
        batch.set_frames(first)
        batch.compute_fill()                                   # mode of each video's first frame (:47)
        missing = np.array([loc is None for loc in locs])
        ij0 = np.empty((nv, 2), np.int32)
        if missing.any():                                      # auto-detect window size .÷ 4 (:102-104)
            batch.set_window((H // 4, W // 4))
            out, _ = batch.step(guess)
            ij0[missing] = out[missing]
            batch.set_window(window_size)
        if (~missing).any():
            out, _ = batch.step(guess)
            ij0[~missing] = out[~missing]
        batch.set_guess(ij0)
        out_all = [ij0[None]]
        if n > 1:
            # decode threads fill a ring of page-locked step-chunks while the GPU tracks the previous chunk
            # (zero-copy footprint reads, the chained kernel: one launch per chunk)
            feeder = FrameFeeder(vids, (H, W), first[0].dtype, n - 1, chunk_steps=chunk_steps, workers=decode_workers)
            try:
                fs = H * W * first[0].dtype.itemsize
                for chunk, Tk in feeder:
                    base = chunk.ctypes.data
                    ptrs = [base + (t * nv + v) * fs for t in range(Tk) for v in range(nv)]
                    ij, _ = batch.track_host_ptrs(ptrs, Tk, W, "footprint")
                    out_all.append(ij)
            finally:
                feeder.close()
        ij = np.concatenate(out_all, axis=0).astype(np.int64)  # (T, nv, 2)
    finally:
        batch.close()
    return ts[:ij.shape[0]], ij
