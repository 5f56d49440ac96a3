import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: minutes of CPU oracle time (still part of -m gpu; deselect with -m 'gpu and not slow')")


def _ensure_built():
    pkg = os.path.join(ROOT, "amg-ann_b200")
    need = [os.path.join(pkg, "libamgb_gen.so"), os.path.join(pkg, "libamgb.so"),
            os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


_ensure_built()


@pytest.fixture(scope="session")
def ab():
    import amg_ann_b200
    return amg_ann_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import binding
    return binding


@pytest.fixture(scope="session")
def gpu_ctx(ab):
    """One context for the whole GPU session.  Fails (does not skip) if the
    CUDA library cannot create a context: there is no CPU fallback to test."""
    ctx = ab.Context(0)
    yield ctx
    ctx.close()
