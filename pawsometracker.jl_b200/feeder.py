"""Host decode → page-locked ring feeder.

The reference decodes on the host (`ffmpeg … -f matroska -` piped into VideoIO, src/PawsomeTracker.jl:155-159)
and `read!`s every frame into the tracker's buffer (:166), one video at a time.  Once the filter runs at tens of
millions of windows per second the decoder is the end-to-end bottleneck, so the batched entry point decodes many
videos concurrently (one worker thread per video slice; OpenCV/FFmpeg releases the GIL) straight into a ring of
page-locked step-chunks, and the GPU tracks chunk k (zero-copy footprint reads, one launch per chunk) while the
workers decode chunk k+1.  Decode itself stays on the host, as BASELINE.json's north_star prescribes.
"""
from __future__ import annotations

import ctypes as C
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from ._lib import check, lib


class PinnedArray:
    """A numpy view over cudaHostAlloc'ed memory (pt_host_alloc); freed by close()."""

    def __init__(self, shape, dtype=np.uint8):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib.pt_host_alloc(nbytes, C.byref(p)))
        self.ptr = p.value
        buf = (C.c_ubyte * nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def close(self):
        if self.ptr:
            self.array = None
            check(lib.pt_host_free(C.c_void_p(self.ptr)))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FrameFeeder:
    """Iterate over step-chunks `(frames[T_k, n, H, W], T_k)` decoded in the background into pinned memory.

    readers: objects with `.eof()` and `.read(out=array)` (api._Resampled) — one per video; every chunk holds the
    next T_k frames of every video, T_k ≤ chunk_steps, and the stream ends when any video does or after max_steps
    frames (`while !eof(vid) && last_frame < n`, :162).  The array yielded stays valid until the next iteration.
    """

    def __init__(self, readers, frame_shape, dtype, max_steps, chunk_steps=None, depth=2, workers=None,
                 budget_bytes=1 << 30):
        self.readers = list(readers)
        self.n = len(self.readers)
        H, W = int(frame_shape[0]), int(frame_shape[1])
        per_step = self.n * H * W * np.dtype(dtype).itemsize
        if chunk_steps is None:
            chunk_steps = max(1, min(32, budget_bytes // max(1, per_step)))
        self.chunk_steps = int(max(1, min(chunk_steps, max(1, max_steps))))
        self.max_steps = int(max_steps)
        self.bufs = [PinnedArray((self.chunk_steps, self.n, H, W), dtype) for _ in range(max(2, depth))]
        self._free = queue.Queue()
        for i in range(len(self.bufs)):
            self._free.put(i)
        self._ready = queue.Queue()
        self._pool = ThreadPoolExecutor(max_workers=workers or min(16, max(1, self.n)))
        self._stop = False
        self._thread = threading.Thread(target=self._produce, daemon=True)
        self._thread.start()

    def _fill_video(self, buf, v, T):
        r = self.readers[v]
        got = 0
        for t in range(T):
            if r.eof():
                break
            r.read(out=buf[t, v])
            got += 1
        return got

    def _produce(self):
        done = 0
        try:
            while not self._stop and done < self.max_steps:
                i = self._free.get()
                if i is None:
                    break
                T = min(self.chunk_steps, self.max_steps - done)
                buf = self.bufs[i].array
                got = list(self._pool.map(lambda v: self._fill_video(buf, v, T), range(self.n)))
                Tk = min(got) if got else 0
                if Tk == 0:
                    break
                self._ready.put((i, Tk))
                done += Tk
                if Tk < T:                     # a video ended inside this chunk
                    break
        except BaseException as e:             # surface decode errors in the consumer
            self._ready.put(e)
            return
        self._ready.put(None)

    def __iter__(self):
        held = None
        while True:
            item = self._ready.get()
            if held is not None:
                self._free.put(held)
                held = None
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            held, Tk = item
            yield self.bufs[held].array[:Tk], Tk

    def close(self):
        self._stop = True
        self._free.put(None)
        self._thread.join(timeout=30)
        self._pool.shutdown(wait=True)
        for b in self.bufs:
            b.close()
