import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def native():
    """Builds (when stale) the CPU-side native pieces the tests need."""
    from sid_b200 import build
    import build_checkers
    build.build_generator()
    build_checkers.build_hostcheck()
    if not os.path.exists(os.path.join(ROOT, "oracle", "build", "liboracle.so")) or os.path.isdir("/root/reference"):
        build_checkers.build_oracle()
    return True


@pytest.fixture(scope="session")
def gpu_ctx():
    import sid_b200
    ctx = sid_b200.Context(max_chunk_bytes=8 << 20)      # small chunks: exercises the chunked host path
    yield ctx
    ctx.close()
