import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests are skipped automatically where no CUDA device exists, so a plain `pytest tests/`
    # in the dev container never tries to launch a kernel.
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref_cpu():
    """The compiled, unmodified reference CPU segment trees (oracle/_ref/cpu) or skip."""
    from oracle.ref_loader import reference_ext

    ext = reference_ext("cpu")
    if ext is None:
        pytest.skip("oracle/_ref/cpu/_torchrl.so not built and /root/reference absent")
    return ext


@pytest.fixture(scope="session")
def ref_funcs():
    """The reference's Python GAE functionals (dev container only)."""
    from oracle.ref_loader import reference_functionals

    F = reference_functionals()
    if F is None:
        pytest.skip("/root/reference not present (GPU box)")
    return F


@pytest.fixture(scope="session")
def ref_samplers():
    """The reference's samplers module, unmodified, under stub imports (dev container only)."""
    from oracle.ref_loader import reference_samplers

    R = reference_samplers()
    if R is None:
        pytest.skip("/root/reference not present (GPU box)")
    return R


@pytest.fixture()
def emul():
    """Install the oracle-backed emulator as rl_b200's backend for one CPU host-logic test."""
    from rl_b200 import ops

    from _emul import OracleBackend

    ops.set_backend(OracleBackend())
    try:
        yield
    finally:
        ops.set_backend(None)


@pytest.fixture(scope="session")
def cuda_backend():
    """The real library on a real GPU (-m gpu tests)."""
    import torch

    from rl_b200 import ops

    assert torch.cuda.is_available()
    ops.set_backend(None)
    return ops.backend()
