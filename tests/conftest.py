import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    # a fresh checkout has no built library (it is git-ignored): build it once, like `__graft_entry__.build()` does
    from quansino_b200 import _lib

    if not _lib.LIB_PATH.exists():
        from quansino_b200.build import build

        build()


@pytest.fixture(scope="session")
def cuda_lib():
    """The C-ABI CUDA library; GPU tests fail (not skip) when it is missing."""
    from quansino_b200 import _lib

    return _lib.load()


@pytest.fixture(params=["warp", "block", "rows"])
def kernel_variant(request, monkeypatch):
    """Run a sweep test through the organisations of the sweep kernel: one warp per replica with a brute-force scan
    (csrc/sweep_kernel.cuh), one CTA per replica (csrc/sweep_block.cuh), and one warp per replica reading per-atom neighbour
    rows (csrc/sweep_rows.cuh; it only engages for single-atom displacement / cell-move tables, other tables run as "auto").
    Without the variable the library picks by table and batch size."""
    monkeypatch.setenv("QMC_FORCE_KERNEL", request.param)
    return request.param
