"""CPU oracle for the PawsomeTracker DoG-window + argmax hot path.

TEST INFRASTRUCTURE ONLY — see oracle/dog_oracle.c header.  PARITY UNPINNED:
the reference (Julia + un-vendored, unpinned ImageFiltering.jl) cannot be run
in this image and ships no golden vectors for this boundary.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package.
"""
from .oracle import (  # noqa: F401
    Oracle,
    OracleResult,
    build,
    numpy_dense_response,
    numpy_factors,
    numpy_mode_u8,
)
