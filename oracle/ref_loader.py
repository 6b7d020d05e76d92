"""TEST INFRASTRUCTURE -- loaders for the real reference (never imported by the product).

``reference_ext(kind)``      the UNMODIFIED reference segment tree compiled by build_ref.py into
                             oracle/_ref/{cpu,cuda}/_torchrl.so (travels to the GPU box).
``reference_functionals()``  the reference's Python value functionals imported by path under a
                             ten-line ``tensordict`` shim -- DEV CONTAINER ONLY (/root/reference is
                             absent on the GPU box; callers must skip when this returns None).
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

_REF_DIR = Path(__file__).resolve().parent / "_ref"
_EXT = {}
_FUNCS = None


def reference_ext(kind: str = "cpu"):
    """Returns the compiled reference module (SumSegmentTreeFp32, ... [CudaSumSegmentTreeFp32, ...]) or None."""
    if kind in _EXT:
        return _EXT[kind]
    so = _REF_DIR / kind / "_torchrl.so"
    if not so.exists():
        try:
            from .build_ref import build

            build(with_cuda=(kind == "cuda"))
        except Exception:
            _EXT[kind] = None
            return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    name = "_torchrl"
    # both variants register the pybind module `_torchrl`; only one can live in a process
    if name in sys.modules and getattr(sys.modules[name], "__file__", None) != str(so):
        other = sys.modules[name]
        _EXT[kind] = other if (kind == "cpu" or hasattr(other, "CudaSumSegmentTreeFp32")) else None
        return _EXT[kind]
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    _EXT[kind] = mod
    return mod


def reference_trees(kind: str = "cpu"):
    """tree_factory for OraclePrioritizedSampler backed by the compiled reference trees."""
    ext = reference_ext(kind)
    if ext is None:
        return None
    return lambda size, is_min: (ext.MinSegmentTreeFp32 if is_min else ext.SumSegmentTreeFp32)(size)


def reference_functionals(root: str = "/root/reference"):
    """Import torchrl/objectives/value/{utils,functional}.py by path."""
    global _FUNCS
    if _FUNCS is not None:
        return _FUNCS
    base = Path(root) / "torchrl" / "objectives" / "value"
    if not base.exists():
        return None
    if "tensordict" not in sys.modules:
        td, tdu = types.ModuleType("tensordict"), types.ModuleType("tensordict.utils")

        class TensorDictBase:  # isinstance() target only
            pass

        def expand_right(t, shape):
            while t.ndim < len(shape):
                t = t.unsqueeze(-1)
            return t.expand(shape)

        td.TensorDictBase, tdu.expand_right, td.utils = TensorDictBase, expand_right, tdu
        sys.modules.update({"tensordict": td, "tensordict.utils": tdu})
    for n in ("torchrl", "torchrl.objectives", "torchrl.objectives.value"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m

    def _load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    _load("torchrl.objectives.value.utils", base / "utils.py")
    _FUNCS = _load("torchrl.objectives.value.functional", base / "functional.py")
    return _FUNCS
