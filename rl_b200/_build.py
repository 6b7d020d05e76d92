"""In-tree build of librlb200.so with nvcc for sm_100a (no torch / pybind linkage)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
SO = PKG / "librlb200.so"
SOURCES = ["abi.cu", "tree.cu", "gae.cu", "gather.cu", "shard.cu", "slice.cu", "framestack.cu"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: librlb200.so cannot be built (set NVCC=/path/to/nvcc)")


def needs_build() -> bool:
    if not SO.exists():
        return True
    t = SO.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "rlb200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -shared -> rl_b200/librlb200.so"""
    if not force and not needs_build():
        return SO
    tmp = SO.with_name(SO.name + ".partial")   # linked beside the target, then renamed: never a half-written library
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-I", str(ROOT / "include"), "-shared", "-o", str(tmp)]
    cmd += [str(CSRC / s) for s in SOURCES] + ["-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    tmp.replace(SO)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
