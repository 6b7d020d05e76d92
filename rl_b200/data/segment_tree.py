"""HBM-resident segment trees with the interface of ``torchrl._torchrl``'s pybind classes.

Drop-in for ``{Sum,Min}SegmentTreeFp{32,64}`` / ``Cuda{Sum,Min}SegmentTreeFp{32,64}``
(reference: torchrl/csrc/pybind.cpp:21-34, segment_tree.h:312-386, cuda_segment_tree.h:246-363):
constructor ``(size[, device])``; properties ``size``, ``capacity``, ``identity_element``, ``device``;
``len()``; ``tree[index]`` / ``at``; ``tree[index] = value`` / ``update``; ``query(l, r)``;
``scan_lower_bound(value)`` (sum tree only); pickling through the leaves.

The values live in one torch tensor of ``2*capacity`` elements on the device, laid out exactly like the
reference's ``values_`` (implicit heap, leaf ``i`` at ``capacity + i``); all arithmetic is done by the
kernels in ``rl_b200/csrc/tree.cu`` through ``rl_b200.ops``.  Tensor overloads never synchronise; the
Python-scalar overloads return Python numbers and therefore do (as the reference's ``.item()`` calls,
cuda_segment_tree.h:50-52,129-137).
"""
from __future__ import annotations

import numbers

import numpy as np
import torch

from .. import ops


class _DeviceSegmentTree:
    _is_min = False
    _dtype = torch.float32

    def __init__(self, size: int, device="cuda", *, out: torch.Tensor | None = None):
        size = int(size)
        if size <= 0:
            raise ValueError("segment tree size must be positive")
        self._size = size
        self._device = torch.device(device)
        be = ops.backend()
        self._capacity = be.tree_capacity(size)
        self._values = be.tree_new(size, self._is_min, self._dtype, self._device, out=out)
        self._workspace = None  # persistent update scratch (ticket, sibling tile, stamps), allocated on first update
        self._epoch = 0

    # ---- properties --------------------------------------------------------------------------------
    @property
    def size(self) -> int:
        return self._size

    @property
    def capacity(self) -> int:
        return self._capacity

    @property
    def identity_element(self) -> float:
        return float(torch.finfo(self._dtype).max) if self._is_min else 0.0

    @property
    def device(self) -> torch.device:
        return self._values.device

    @property
    def values(self) -> torch.Tensor:
        """The raw ``2*capacity`` heap tensor (same layout as the reference's ``values_``)."""
        return self._values

    def __len__(self) -> int:
        return self._size

    # ---- helpers -----------------------------------------------------------------------------------
    def _as_index(self, index) -> torch.Tensor:
        if isinstance(index, torch.Tensor):
            if index.dtype != torch.int64:
                raise RuntimeError("index must be an int64 tensor")
            return index.to(self.device)
        return torch.as_tensor(np.asarray(index, dtype=np.int64), device=self.device)

    def _as_value(self, value) -> torch.Tensor:
        if isinstance(value, torch.Tensor):
            if value.dtype != self._dtype:
                raise RuntimeError("value dtype must match the tree dtype")
            return value.to(self.device)
        return torch.as_tensor(np.asarray(value), dtype=self._dtype, device=self.device)

    def _next_epoch(self) -> int:
        self._epoch += 1
        if self._epoch >= 0xFFFFFFFF:  # wrap-around: stamps from 2^32 calls ago could alias
            if self._workspace is not None:
                self._workspace.zero_()
            self._epoch = 1
        return self._epoch

    def _ensure_workspace(self, n: int):
        if self._workspace is None:
            self._workspace = ops.backend().tree_workspace(self._size, self.device)
        return self._workspace

    # ---- At / __getitem__  (segment_tree.h:56-79) ----------------------------------------------------
    def at(self, index):
        if isinstance(index, numbers.Integral):
            return self.at(torch.tensor([int(index)], dtype=torch.int64)).item()
        idx = self._as_index(index)
        out = ops.backend().tree_at(self._values, self._capacity, idx.reshape(-1)).reshape(idx.shape)
        if isinstance(index, torch.Tensor):
            return out
        return out.cpu().numpy()

    __getitem__ = at

    # ---- Update / __setitem__  (segment_tree.h:83-139) ----------------------------------------------
    def update(self, index, value) -> None:
        idx = self._as_index(index).reshape(-1)
        val = self._as_value(value).reshape(-1)
        if val.numel() != 1 and val.numel() != idx.numel():
            raise RuntimeError("value must have one element or the same number of elements as index")
        sum_t, min_t = (None, self._values) if self._is_min else (self._values, None)
        ops.backend().tree_update(sum_t, min_t, self._capacity, idx, val, self._ensure_workspace(idx.numel()),
                                  self._next_epoch())

    __setitem__ = update

    # ---- Query  (segment_tree.h:143-162) --------------------------------------------------------------
    def query(self, l, r, *, root_fast_path: bool = True):
        scalar = isinstance(l, numbers.Integral) and isinstance(r, numbers.Integral)
        if scalar:
            if not l < r:
                raise ValueError("query needs l < r")  # assert(l < r), segment_tree.h:144
        li, ri = self._as_index(l), self._as_index(r)
        if li.shape != ri.shape:
            raise RuntimeError("l and r must have the same shape")
        out = ops.backend().tree_query(self._values, self._size, self._capacity, self._is_min, li.reshape(-1),
                                       ri.reshape(-1), root_fast_path).reshape(li.shape)
        if scalar:
            return out.item()
        if isinstance(l, torch.Tensor):
            return out
        return out.cpu().numpy()

    # ---- pickling: leaves + device, rebuilt on load (segment_tree.h:375-385, cuda_segment_tree.h:286-305)
    def dump_leaves(self) -> torch.Tensor:
        """DumpValues: the ``size`` leaves, as a new device tensor (one D2D copy, no per-element calls)."""
        return self._values[self._capacity:self._capacity + self._size].clone()

    def load_leaves(self, leaves) -> None:
        """LoadValues: overwrite the leaves and rebuild every internal node bottom-up."""
        leaves = torch.as_tensor(leaves, dtype=self._dtype).to(self.device).reshape(-1)
        if leaves.numel() != self._size:
            raise RuntimeError(f"expected {self._size} leaves, got {leaves.numel()}")
        self._values[self._capacity:self._capacity + self._size].copy_(leaves)
        ops.backend().tree_rebuild(self._values, self._capacity, self._is_min)

    def __getstate__(self):
        return {"size": self._size, "device": str(self._device), "leaves": self.dump_leaves().cpu().numpy()}

    def __setstate__(self, state):
        self.__init__(state["size"], state["device"])
        self.load_leaves(torch.from_numpy(state["leaves"]))

    def __deepcopy__(self, memo):
        new = type(self)(self._size, self.device)
        new._values.copy_(self._values)
        return new

    def __repr__(self) -> str:
        return f"{type(self).__name__}(size={self._size}, capacity={self._capacity}, device={self.device})"


class _SumMixin:
    # ---- ScanLowerBound  (segment_tree.h:249-264) -----------------------------------------------------
    def scan_lower_bound(self, value):
        if isinstance(value, numbers.Real):
            return int(self.scan_lower_bound(torch.tensor([value], dtype=self._dtype)).item())
        val = self._as_value(value)
        out = ops.backend().tree_scan_lower_bound(self._values, self._size, self._capacity,
                                                  val.reshape(-1)).reshape(val.shape)
        if isinstance(value, torch.Tensor):
            return out
        return out.cpu().numpy()


class SumSegmentTreeFp32(_SumMixin, _DeviceSegmentTree):
    _is_min, _dtype = False, torch.float32


class SumSegmentTreeFp64(_SumMixin, _DeviceSegmentTree):
    _is_min, _dtype = False, torch.float64


class MinSegmentTreeFp32(_DeviceSegmentTree):
    _is_min, _dtype = True, torch.float32


class MinSegmentTreeFp64(_DeviceSegmentTree):
    _is_min, _dtype = True, torch.float64


# the reference exposes the device variants under these names (csrc/pybind.cpp:27-33)
CudaSumSegmentTreeFp32, CudaSumSegmentTreeFp64 = SumSegmentTreeFp32, SumSegmentTreeFp64
CudaMinSegmentTreeFp32, CudaMinSegmentTreeFp64 = MinSegmentTreeFp32, MinSegmentTreeFp64
