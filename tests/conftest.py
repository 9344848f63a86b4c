import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand with gcc."""
    from oracle import Oracle, build
    build()
    return Oracle()


@pytest.fixture(scope="session")
def synth():
    """The reference's synthetic-video test recipe (tools/synth.py): test infrastructure, not part of the package."""
    from tools import synth as m
    return m


@pytest.fixture(scope="session")
def pkg():
    """The product package. Importing it requires the built libpawsome_cuda.so."""
    so = os.path.join(ROOT, "pawsometracker.jl_b200", "libpawsome_cuda.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    import pt_import
    return pt_import.load()


@pytest.fixture(scope="session")
def gpu_pkg(pkg):
    n = pkg.lib.pt_device_count()
    if n < 1:
        pytest.fail("no CUDA device visible: the product path has no CPU fallback "
                    f"(pt_device_count={n}: {pkg._lib.last_error()})")
    return pkg
