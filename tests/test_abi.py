"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/rlb200.h declares.
No compute is launched here (CPU suite)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def so():
    from rl_b200 import _build

    return _build.build()


def _declared():
    text = (ROOT / "include" / "rlb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rlb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(so):
    L = ctypes.CDLL(str(so))
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/rlb200.h but not exported by {so.name}"


def test_python_prototypes_cover_header(so):
    from rl_b200 import ops

    assert ops.exported_symbols() == _declared()
    L = ops.load_library()
    assert L.rlb_version() == 100


def test_host_only_entry_points(so):
    from rl_b200 import ops

    L = ops.load_library()
    # capacity rule: smallest power of two STRICTLY greater than size (csrc/segment_tree.h:44-48)
    for size, cap in [(1, 2), (2, 4), (15, 16), (16, 32), (1000, 1024), (1024, 2048), (1_000_000, 1 << 20),
                      (1_250_000, 1 << 21), (6_250_000, 1 << 23)]:
        assert L.rlb_tree_capacity(size) == cap
    assert L.rlb_tree_update_workspace_bytes(1000) == (1 << 20) + 1024 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(so):
    from rl_b200 import ops

    ops.set_backend(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.backend()
    L = ops.load_library()
    assert L.rlb_device_sm_count() < 0
    assert L.rlb_last_error()  # a CUDA error string, not a silent fallback
    # argument validation happens before any launch
    assert L.rlb_gae(None, None, None, None, None, 0.99, 0.94, 4, 8, 1, 0, None, None, None) == -1
    assert b"null pointer" in L.rlb_last_error()


def test_sass_uses_bulk_copy_engine(so):
    """The gather kernel's DMA role must compile to UBLKCP (cp.async.bulk) + mbarrier SYNCS."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(so)], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", str(so)], capture_output=True, text=True).stdout
