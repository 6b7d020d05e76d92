"""Writers: the host-side mirror of ``torchrl.data.replay_buffers.writers`` (round-robin family).

    Writer                       abstract contract                   writers.py:43-118
    RoundRobinWriter             cursor = (cursor + n) % max_size     writers.py:148-310
    TensorDictRoundRobinWriter   + stamps data["index"]               writers.py:313-361

A writer decides WHERE data goes; the bytes are moved by the storage.  After every write it calls
``mark_update(index)`` on each buffer attached to the storage (writers.py:232-235) so prioritized
samplers give the new items their default priority.

One scheduling difference from the reference: when a batch does not wrap around the end of the storage
the cursor handed to ``storage.set`` is a ``slice`` (one contiguous copy per leaf, which for host data is
a direct H2D DMA into place) instead of an index tensor (an ``index_put_`` per leaf); the returned /
stamped index tensor is the same ``arange(cursor, cursor+n) % max_size``.
"""
from __future__ import annotations

import abc
import json
from copy import copy
from pathlib import Path
from typing import Any

import torch
from torch.utils import _pytree as pytree

from .storages import Storage
from .tensordict_lite import expand_as_right, is_tensor_collection
from .utils import _is_int


class Writer(abc.ABC):
    """A ReplayBuffer base Writer class (writers.py:43-118)."""

    _storage: Storage
    _rng: torch.Generator | None = None

    def __init__(self, compilable: bool = False) -> None:
        self._storage = None
        self._compilable = compilable

    def register_storage(self, storage: Storage) -> None:
        self._storage = storage

    @abc.abstractmethod
    def add(self, data: Any) -> int:
        """Inserts one piece of data at an appropriate index, and returns that index."""

    @abc.abstractmethod
    def extend(self, data) -> torch.Tensor:
        """Inserts a series of data points at appropriate indices, and returns a tensor containing the indices."""

    @abc.abstractmethod
    def _empty(self, empty_write_count: bool = True) -> None:
        ...

    @abc.abstractmethod
    def dumps(self, path) -> None:
        ...

    @abc.abstractmethod
    def loads(self, path) -> None:
        ...

    @abc.abstractmethod
    def state_dict(self) -> dict:
        ...

    @abc.abstractmethod
    def load_state_dict(self, state_dict: dict) -> None:
        ...

    def _replicate_index(self, index):
        # for multi-dim storages every written item is addressed by a full coordinate (writers.py:85-110)
        if self._storage.ndim == 1:
            return index
        device = index.device if isinstance(index, torch.Tensor) else torch.device("cpu")
        mesh = torch.stack(
            torch.meshgrid(*(torch.arange(d, device=device) for d in self._storage.shape[1:]), indexing="ij"),
            -1).flatten(0, -2)
        if _is_int(index):
            first = torch.as_tensor(int(index), device=device).expand(mesh.shape[0], 1)
            return torch.cat([first, mesh], 1)
        return torch.cat([index.repeat_interleave(mesh.shape[0]).unsqueeze(1), mesh.repeat(index.numel(), 1)], 1)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"

    def __getstate__(self):
        state = copy(self.__dict__)
        state["_rng"] = None
        return state


class RoundRobinWriter(Writer):
    """Writes at ``cursor`` and advances it modulo the storage size (writers.py:148-310)."""

    def __init__(self, compilable: bool = False) -> None:
        super().__init__(compilable=compilable)
        self._cursor = 0
        self._write_count = 0

    def dumps(self, path) -> None:
        path = Path(path).absolute()
        path.mkdir(exist_ok=True, parents=True)
        with open(path / "metadata.json", "w") as file:
            json.dump({"cursor": self._cursor}, file)

    def loads(self, path) -> None:
        with open(Path(path).absolute() / "metadata.json") as file:
            self._cursor = json.load(file)["cursor"]

    def state_dict(self) -> dict:
        return {"_cursor": self._cursor}

    def load_state_dict(self, state_dict: dict) -> None:
        self._cursor = state_dict["_cursor"]

    def _empty(self, empty_write_count: bool = True) -> None:
        self._cursor = 0
        if empty_write_count:
            self._write_count = 0

    def _mark_update_entities(self, index) -> None:
        for ent in self._storage._attached_entities_iter():
            ent.mark_update(index)

    # ---- shared mechanics
    def _stamp(self, data, index) -> None:
        """Hook for TensorDictRoundRobinWriter: record where each item went inside the data itself."""

    def add(self, data: Any):
        cur = self._cursor
        # the cursor moves first, as in the reference, so concurrent writers never collide (writers.py:175-181)
        self._cursor = (cur + 1) % self._storage._max_size_along_dim0(single_data=data)
        self._write_count += 1
        self._stamp(data, cur)
        self._storage.set(cur, data)
        index = self._replicate_index(cur)
        self._mark_update_entities(index)
        return index

    def extend(self, data) -> torch.Tensor:
        cur = self._cursor
        if is_tensor_collection(data) or isinstance(data, (torch.Tensor, list)):
            n = len(data)
        else:
            n = len(pytree.tree_leaves(data)[0])
        if n == 0:
            raise RuntimeError(f"Expected at least one element in extend. Got {data=}")
        device = data.device if hasattr(data, "device") else None
        max0 = self._storage._max_size_along_dim0(batched_data=data)
        index = torch.arange(cur, cur + n, dtype=torch.long, device=device)
        if cur + n > max0:
            index = index % max0
        self._cursor = (cur + n) % max0
        self._write_count += n
        self._stamp(data, index)
        if self._extend_fused(cur, n, max0, data, index):
            return index
        self._storage.set(slice(cur, cur + n) if cur + n <= max0 else index, data)
        index = self._replicate_index(index)
        self._mark_update_entities(index)
        return index

    def _extend_fused(self, cur: int, n: int, max0: int, data, index) -> bool:
        """Rows and default priorities of the batch in one launch (``TensorStorage._extend_range``): a writer batch is
        always the modular slot range (cur + arange(n)) % max0, for which neither the row write nor the tree update
        needs the index tensor.  The first attached buffer whose sampler accepts the range form rides in the row
        kernel; every other attached buffer is told through ``mark_update`` as usual."""
        storage = self._storage
        if not hasattr(storage, "_extend_range") or not storage._fits_range(n, data):
            return False
        ents = list(storage._attached_entities_iter())
        trees, fused_ent = None, None
        for ent in ents:
            make = getattr(getattr(ent, "_sampler", None), "_range_update", None)
            if make is not None:
                trees = make(max0, storage=storage)
                if trees is not None:
                    fused_ent = ent
                    break
        storage._extend_range(cur, n, data, trees)
        for ent in ents:
            if ent is not fused_ent:
                ent.mark_update(index)
        return True

    def __repr__(self) -> str:
        full = self._storage._is_full if self._storage is not None else None
        return f"{self.__class__.__name__}(cursor={int(self._cursor)}, full_storage={full})"


class TensorDictRoundRobinWriter(RoundRobinWriter):
    """A RoundRobin Writer for tensordict-based buffers: also sets ``data["index"]`` (writers.py:313-361)."""

    def _stamp(self, data, index) -> None:
        if not is_tensor_collection(data):
            return
        idx = torch.as_tensor(index, dtype=torch.long, device=data.device)
        data.set("index", expand_as_right(idx, data))
