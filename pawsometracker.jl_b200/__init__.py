"""pawsometracker.jl_b200 — B200-native DoG-window + argmax hot path of
PawsomeTracker.jl behind the reference's own Tracker / track() surface.

The directory name contains a dot, so import it through the repo-root helper:

    import pt_import; pkg = pt_import.load()       # → module `pawsometracker_jl_b200`

Everything numerical runs in libpawsome_cuda.so (hand-written sm_100a CUDA,
C ABI in include/pawsome.h); importing this package fails loudly when that
library has not been built — there is no CPU fallback.
"""
from ._lib import LIB_PATH, PT_PIX_F32, PT_PIX_U8, PawsomeError, lib  # noqa: F401
from .api import (  # noqa: F401
    DEFAULT_MAX_DURATION_SECONDS,
    ArrayVideo,
    CartesianIndex,
    CvVideo,
    Diagnose,
    get_guess,
    get_start_ij_and_tracker,
    split_chains,
    track,
    track_batch,
    track_one,
    track_segments,
)
from .feeder import FrameFeeder, PinnedArray  # noqa: F401
from .tracker import (  # noqa: F401
    Tracker,
    TrackerBatch,
    factors_f32,
    fix_window_size,
    get_sigma,
    guess_window_size,
    kernel_len,
)

__version__ = "0.1.0"
