"""TEST INFRASTRUCTURE -- CPU oracle for the replay-and-advantage hot path.

Nothing in the product package ``rl_b200`` imports this package.  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.

Contents
  rlb_oracle.c   plain-C restatement of the reference segment tree, PER sample arithmetic, GAE
                 loop and row gather (each function cites the reference file:line).
  per_oracle.py  the Python-level glue of PrioritizedSampler (sample / update_priority /
                 default_priority / mark_update) restated on top of either the C oracle tree or
                 the compiled reference tree.
  gae_torch.py   the reference's two CPU GAE code paths restated as the same torch-op sequences
                 (time loop and pad+conv1d), used as the CPU baseline where /root/reference is
                 absent (the GPU box).
  build_ref.py   compiles the UNMODIFIED reference csrc into oracle/_ref/ (git-ignored).
  ref_loader.py  loads oracle/_ref and (dev container only) the reference's Python functionals.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_LIB = None


def build() -> Path:
    """Compile rlb_oracle.c -> liborc.so (gcc, ~1 s).  Building the checker is not using it."""
    so = _HERE / "liborc.so"
    src = _HERE / "rlb_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(_HERE), "liborc.so"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    L = ctypes.CDLL(str(build()))
    c_i64, c_f32, c_f64, c_int, vp = (ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_int,
                                      ctypes.c_void_p)
    L.orc_tree_new.restype = vp
    L.orc_tree_new.argtypes = [c_i64, c_int]
    L.orc_tree_free.argtypes = [vp]
    L.orc_tree_capacity.restype = c_i64
    L.orc_tree_capacity.argtypes = [vp]
    L.orc_tree_size.restype = c_i64
    L.orc_tree_size.argtypes = [vp]
    L.orc_tree_values.restype = vp
    L.orc_tree_values.argtypes = [vp]
    L.orc_tree_update.argtypes = [vp, vp, vp, c_i64, c_int]
    L.orc_tree_at.argtypes = [vp, vp, vp, c_i64]
    L.orc_tree_query.restype = c_f32
    L.orc_tree_query.argtypes = [vp, c_i64, c_i64]
    L.orc_tree_query_walk.restype = c_f32
    L.orc_tree_query_walk.argtypes = [vp, c_i64, c_i64]
    L.orc_tree_scan_lower_bound.argtypes = [vp, vp, vp, c_i64]
    L.orc_tree_dump_leaves.argtypes = [vp, vp]
    L.orc_tree_load_leaves.argtypes = [vp, vp]
    L.orc_per_sample.restype = c_int
    L.orc_per_sample.argtypes = [vp, vp, c_i64, vp, c_i64, c_int, c_int, vp, vp, vp, vp]
    L.orc_gae_f32.argtypes = [vp, vp, vp, vp, vp, c_f32, c_f32, c_i64, c_i64, c_i64, vp, vp]
    L.orc_gae_f64.argtypes = [vp, vp, vp, vp, vp, c_f64, c_f64, c_i64, c_i64, c_i64, vp, vp]
    L.orc_td_lambda_f32.argtypes = [vp, vp, vp, vp, c_f32, c_f32, c_i64, c_i64, c_i64, vp]
    L.orc_td_lambda_f64.argtypes = [vp, vp, vp, vp, c_f64, c_f64, c_i64, c_i64, c_i64, vp]
    L.orc_affine_scan_f32.argtypes = [vp, vp, c_i64, c_i64, c_i64, vp]
    L.orc_affine_scan_f64.argtypes = [vp, vp, c_i64, c_i64, c_i64, vp]
    L.orc_gather_rows.restype = c_int
    L.orc_gather_rows.argtypes = [vp, c_i64, c_i64, c_i64, vp, c_i64, vp]
    _LIB = L
    return L
