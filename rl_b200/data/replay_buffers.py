"""Replay buffers: the host-side mirror of ``torchrl.data.replay_buffers.replay_buffers``.

    ReplayBuffer                        composable storage + sampler + writer    replay_buffers.py:109-1640
    PrioritizedReplayBuffer             PER shortcut                             replay_buffers.py:1393-1640
    TensorDictReplayBuffer              packs "index"/info into the sample       replay_buffers.py:1644-2022
    TensorDictPrioritizedReplayBuffer   PER + priority_key write-back            replay_buffers.py:2025-2230

Same constructor keywords, method names, locking (one RLock around sampler + storage access), info
packing and error messages for the calls on the hot path: ``add / extend / sample / update_priority /
update_tensordict_priority / mark_update / empty / set_rng / state_dict / dumps / loads``.  Orchestration
features that are not on the path (transforms, shared-memory multiprocessing, ensembles,
remote / Ray buffers) are out of scope -- see DESIGN.md.

With a CUDA ``LazyTensorStorage`` and a ``PrioritizedSampler`` a ``sample()`` is three launches --
``torch.rand``, ``rlb_per_sample``, ``rlb_gather`` -- and no host synchronisation; the sampled indices
never leave the device.
"""
from __future__ import annotations

import collections
import contextlib
import json
import threading
import warnings
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Any, Callable

import torch

from .samplers import PrioritizedSampler, RandomSampler, Sampler
from .storages import ListStorage, Storage
from .tensordict_lite import expand_as_right, is_tensor_collection, stack_tds
from .utils import INT_CLASSES, _is_int, _reduce, _to_torch
from .writers import RoundRobinWriter, TensorDictRoundRobinWriter, Writer


def _storage_index(index, storage: Storage):
    """Move a sampled index to the storage's device if needed (replay_buffers.py:83-96)."""
    dev = getattr(storage, "device", None)
    if dev is None or dev == "auto":
        return index
    dev = torch.device(dev)

    def move(ix):
        return ix.to(dev) if isinstance(ix, torch.Tensor) and ix.device != dev else ix

    return tuple(move(i) for i in index) if isinstance(index, tuple) else move(index)


def _stack_anything(items):
    """Default collate for ListStorage (storages.py:2676-2685)."""
    if isinstance(items, (list, tuple)) and len(items) and is_tensor_collection(items[0]):
        return stack_tds(list(items))
    if isinstance(items, (list, tuple)) and len(items) and isinstance(items[0], torch.Tensor):
        return torch.stack(list(items))
    if isinstance(items, (list, tuple)) and len(items) and isinstance(items[0], (tuple, list, dict)):
        from torch.utils import _pytree as pytree

        return pytree.tree_map(lambda *xs: torch.stack(xs), *items)
    return items


class ReplayBuffer:
    """A generic, composable replay buffer class (replay_buffers.py:109-292).

    Keyword Args:
        storage (Storage or callable): defaults to ``ListStorage(max_size=1_000)``.
        sampler (Sampler or callable): defaults to :class:`RandomSampler`.
        writer (Writer or callable): defaults to :class:`RoundRobinWriter`.
        collate_fn (callable): merges a list of samples; identity for tensor storages.
        batch_size (int): default batch size of :meth:`sample`.
        dim_extend (int): which dim of the data ``extend`` iterates over (defaults to ``storage.ndim - 1``).
        generator (torch.Generator): random generator shared by storage, sampler and writer.
        prefetch (int, optional): number of next batches prepared by a thread pool (replay_buffers.py:340-341,
            1155-1164).  With an HBM-resident buffer every launch is already asynchronous, so this only hides the
            python / launch latency of ``_sample``; worker threads launch on their own current stream.
        pin_memory, transform, shared, compilable, delayed_init: accepted for signature compatibility; anything but
            the default raises ``NotImplementedError`` (off the hot path).
    """

    def __init__(self, *, storage=None, sampler=None, writer=None, collate_fn: Callable | None = None,
                 pin_memory: bool = False, prefetch: int | None = None, transform=None, transform_factory=None,
                 batch_size: int | None = None, dim_extend: int | None = None, checkpointer=None,
                 generator: torch.Generator | None = None, shared: bool = False, compilable: bool | None = None,
                 delayed_init: bool | None = None) -> None:
        for name, val in (("transform", transform), ("transform_factory", transform_factory),
                          ("shared", shared), ("pin_memory", pin_memory), ("delayed_init", delayed_init)):
            if val:
                raise NotImplementedError(
                    f"ReplayBuffer({name}=...) is orchestration outside the B200 hot path and is not provided.")
        if dim_extend is not None and dim_extend < 0:
            raise ValueError("dim_extend must be a positive value.")
        self._batch_size = batch_size
        self._prefetch = bool(prefetch)
        self._prefetch_cap = prefetch or 0
        self._prefetch_queue = collections.deque()
        self._prefetch_executor = ThreadPoolExecutor(max_workers=self._prefetch_cap) if self._prefetch else None
        self._futures_lock = threading.RLock()
        self._replay_lock = threading.RLock()
        self._write_lock = contextlib.nullcontext()
        self.shared = False

        self._storage = self._make(storage, Storage, lambda: ListStorage(max_size=1_000), "storage")
        self._storage.attach(self)
        self._sampler = self._make(sampler, Sampler, RandomSampler, "sampler")
        self._writer = self._make(writer, Writer, RoundRobinWriter, "writer")
        self._writer.register_storage(self._storage)
        if collate_fn is None:
            collate_fn = _stack_anything if isinstance(self._storage, ListStorage) else (lambda x: x)
        self._collate_fn = collate_fn
        if self._batch_size is None and getattr(self._sampler, "drop_last", False):
            raise ValueError(
                "Samplers with drop_last=True must work with a predictable batch-size. "
                "Please pass the batch-size to the ReplayBuffer constructor.")
        self._dim_extend = dim_extend if dim_extend is not None else self._storage.ndim - 1
        self.set_rng(generator)
        self._initialize_prioritized_sampler()

    @staticmethod
    def _make(obj, cls, default, what):
        if obj is None:
            return default()
        if not isinstance(obj, cls) and callable(obj):
            obj = obj()
        if not isinstance(obj, cls):
            raise TypeError(f"{what} must be either a {cls.__name__} or a callable returning a {what} instance.")
        return obj

    def _initialize_prioritized_sampler(self) -> None:
        # a prioritized sampler attached to a pre-filled storage starts with default priorities everywhere
        # (replay_buffers.py:426-458)
        if isinstance(self._sampler, PrioritizedSampler) and len(self._storage) > 0:
            device = getattr(self._storage, "device", None)
            if device == "auto":
                device = None
            n = len(self._storage)
            indices = torch.arange(n, dtype=torch.long, device=device)
            prio = torch.full((n,), self._sampler.default_priority, dtype=torch.float, device=device)
            self._sampler.update_priority(indices, prio, storage=self._storage)

    # ---- small accessors -----------------------------------------------------------------------------
    def set_rng(self, generator) -> None:
        self._rng = generator
        self._storage._rng = generator
        self._sampler._rng = generator
        self._writer._rng = generator

    @property
    def dim_extend(self) -> int:
        return self._dim_extend

    @property
    def batch_size(self):
        return self._batch_size

    @property
    def storage(self) -> Storage:
        return self._storage

    @property
    def sampler(self) -> Sampler:
        return self._sampler

    @property
    def writer(self) -> Writer:
        return self._writer

    @property
    def write_count(self) -> int:
        return self._writer._write_count

    def __len__(self) -> int:
        with self._replay_lock:
            return len(self._storage)

    def __repr__(self) -> str:
        return (f"{type(self).__name__}(storage={self._storage}, sampler={self._sampler}, writer={self._writer}, "
                f"batch_size={self._batch_size})")

    def __getitem__(self, index):
        if isinstance(index, str) or (isinstance(index, tuple) and all(isinstance(i, str) for i in index)):
            return self[:][index]
        if isinstance(index, tuple):
            if len(index) == 1:
                return self[index[0]]
            return self[:][index]
        index = _to_torch(index) if not isinstance(index, (slice, type(None), type(Ellipsis), int)) else index
        with self._replay_lock:
            data = self._storage[index]
        if not isinstance(index, INT_CLASSES):
            data = self._collate_fn(data)
        return data

    def __setitem__(self, index, value) -> None:
        with self._replay_lock, self._write_lock:
            self._storage[index] = value

    def _transpose(self, data):
        d = self.dim_extend
        if is_tensor_collection(data):
            return data.apply(lambda t: t.transpose(d, 0), batch_size=torch.Size(
                torch.empty(data.batch_size, device="meta").transpose(d, 0).shape))
        from torch.utils import _pytree as pytree

        return pytree.tree_map(lambda x: x.transpose(d, 0), data)

    # ---- writes ------------------------------------------------------------------------------------
    def add(self, data: Any):
        """Add a single element to the replay buffer; returns the index where it lives."""
        if data is None:
            return torch.zeros((0, self._storage.ndim), dtype=torch.long)
        return self._add(data)

    def _add(self, data):
        with self._replay_lock, self._write_lock:
            index = self._writer.add(data)
            self._sampler.add(index)
        return index

    def _extend(self, data, *, update_priority: bool = True) -> torch.Tensor:
        with self._replay_lock, self._write_lock:
            if self.dim_extend > 0:
                data = self._transpose(data)
            index = self._writer.extend(data)
            self._sampler.extend(index)
        return index

    def extend(self, data, *, update_priority: bool | None = None) -> torch.Tensor:
        """Extends the replay buffer with one or more elements contained in an iterable; returns their indices.

        A tuple is a pytree (all leaves share the leading batch dim); a list is a stack of single items.
        """
        if update_priority is not None:
            raise NotImplementedError(
                "update_priority is not supported in this class. See "
                ":meth:`~torchrl.data.TensorDictReplayBuffer.extend` for more details.")
        if data is None:
            return torch.zeros((0, self._storage.ndim), dtype=torch.long)
        return self._extend(data)

    def update_priority(self, index, priority) -> None:
        if isinstance(index, tuple):
            index = torch.stack(index, -1)
        priority = torch.as_tensor(priority)
        if self.dim_extend > 0 and priority.ndim > 1:
            priority = self._transpose(priority).flatten()
        with self._replay_lock, self._write_lock:
            self._sampler.update_priority(index, priority, storage=self.storage)

    def mark_update(self, index) -> None:
        self._sampler.mark_update(index, storage=self._storage)

    def empty(self, empty_write_count: bool = True) -> None:
        """Empties the replay buffer and resets the cursor to 0."""
        self._writer._empty(empty_write_count=empty_write_count)
        self._sampler._empty()
        self._storage._empty()

    # ---- reads -------------------------------------------------------------------------------------
    def _sample(self, batch_size: int) -> tuple[Any, dict]:
        with self._replay_lock, self._write_lock:
            index, info = self._sampler.sample(self._storage, batch_size)
            info["index"] = index
            # indices a sampler of this engine produced are in range by construction: the unchecked gather
            six = _storage_index(index, self._storage)
            trusted = getattr(self._storage, "_get_trusted", None)
            data = trusted(six) if (trusted is not None and isinstance(six, torch.Tensor)) else self._storage.get(six)
        if not isinstance(index, INT_CLASSES):
            data = self._collate_fn(data)
        return data, info

    def sample(self, batch_size: int | None = None, return_info: bool = False) -> Any:
        """Samples a batch of data from the replay buffer (sampler -> indices, storage -> rows).

        Returns the batch, or ``(batch, info)`` when ``return_info`` is set.
        """
        if batch_size is not None and self._batch_size is not None and batch_size != self._batch_size:
            warnings.warn(
                f"Got conflicting batch_sizes in constructor ({self._batch_size}) and `sample` ({batch_size}). "
                "Refer to the ReplayBuffer documentation for a proper usage of the batch-size arguments. "
                "The batch-size provided to the sample method will prevail.")
        elif batch_size is None and self._batch_size is not None:
            batch_size = self._batch_size
        elif batch_size is None:
            raise RuntimeError(
                "batch_size not specified. You can specify the batch_size when constructing the replay buffer, "
                "or pass it to the sample method. Refer to the ReplayBuffer documentation for a proper usage of "
                "the batch-size arguments.")
        if not self._prefetch:
            data, info = self._sample(batch_size)
        else:
            with self._futures_lock:
                while (len(self._prefetch_queue) < min(self._sampler._remaining_batches, self._prefetch_cap)
                       and not self._sampler.ran_out) or not len(self._prefetch_queue):
                    self._prefetch_queue.append(self._prefetch_executor.submit(self._sample, batch_size))
                data, info = self._prefetch_queue.popleft().result()
        if return_info:
            dev = getattr(self.storage, "device", None)
            if dev is not None and dev != "auto":
                info = {k: (tuple(x.to(dev) for x in v) if isinstance(v, tuple) else
                            (v.to(dev) if hasattr(v, "to") else v)) for k, v in info.items()}
            return data, info
        return data

    def __iter__(self):
        if self._sampler.ran_out:
            self._sampler.ran_out = False
        if self._batch_size is None:
            raise RuntimeError("Cannot iterate over the replay buffer. Batch_size was not specified during "
                               "construction of the replay buffer.")
        while not self._sampler.ran_out:
            yield self.sample()

    # ---- (de)serialisation -------------------------------------------------------------------------
    def state_dict(self) -> dict:
        return {"_storage": self._storage.state_dict(), "_sampler": self._sampler.state_dict(),
                "_writer": self._writer.state_dict(), "_batch_size": self._batch_size}

    def load_state_dict(self, state_dict: dict) -> None:
        self._storage.load_state_dict(state_dict["_storage"])
        self._sampler.load_state_dict(state_dict["_sampler"])
        self._writer.load_state_dict(state_dict["_writer"])
        self._batch_size = state_dict["_batch_size"]

    def dumps(self, path) -> None:
        """Saves the replay buffer on disk: ``storage/``, ``sampler/``, ``writer/``, ``buffer_metadata.json``
        (directory layout of replay_buffers.py:856-938)."""
        path = Path(path).absolute()
        path.mkdir(exist_ok=True, parents=True)
        self._storage.dumps(path / "storage")
        self._sampler.dumps(path / "sampler")
        self._writer.dumps(path / "writer")
        with open(path / "buffer_metadata.json", "w") as file:
            json.dump({"batch_size": self._batch_size}, file)

    def loads(self, path) -> None:
        path = Path(path).absolute()
        self._storage.loads(path / "storage")
        self._sampler.loads(path / "sampler")
        self._writer.loads(path / "writer")
        with open(path / "buffer_metadata.json") as file:
            self._batch_size = json.load(file)["batch_size"]

    save = dump = dumps
    load = loads


class PrioritizedReplayBuffer(ReplayBuffer):
    """Prioritized replay buffer: ``ReplayBuffer`` with a :class:`PrioritizedSampler`
    (replay_buffers.py:1393-1640).  All arguments are keyword-only."""

    def __init__(self, *, alpha: float, beta: float, eps: float = 1e-8, dtype: torch.dtype = torch.float,
                 storage=None, sampler_device=None, collate_fn=None, batch_size: int | None = None,
                 dim_extend: int | None = None, generator=None, **kwargs) -> None:
        if storage is None:
            storage = ListStorage(max_size=1_000)
        elif not isinstance(storage, Storage) and callable(storage):
            storage = storage()
        sampler = PrioritizedSampler(storage.max_size, alpha, beta, eps, dtype, device=sampler_device)
        super().__init__(storage=storage, sampler=sampler, collate_fn=collate_fn, batch_size=batch_size,
                         dim_extend=dim_extend, generator=generator, **kwargs)


class TensorDictReplayBuffer(ReplayBuffer):
    """TensorDict-specific wrapper around :class:`ReplayBuffer` (replay_buffers.py:1644-2022): the sampled
    tensordict carries ``"index"`` and the sampler's info (``"priority_weight"``), and priorities can be read
    from ``priority_key`` of a tensordict."""

    def __init__(self, *, priority_key: str = "td_error", **kwargs) -> None:
        if kwargs.get("writer") is None:
            kwargs["writer"] = TensorDictRoundRobinWriter
        super().__init__(**kwargs)
        self.priority_key = priority_key

    def _get_priority_item(self, tensordict):
        priority = tensordict.get(self.priority_key, None)
        if priority is None:
            return self._sampler.default_priority
        if self._storage.ndim > 1:
            priority = priority.flatten(0, self._storage.ndim - 1)
        try:
            priority = _reduce(priority, self._sampler.reduction) if priority.numel() > 1 else priority.item()
        except ValueError:
            raise ValueError(
                f"Found a priority key of size {tensordict.get(self.priority_key).shape} but expected scalar value")
        return priority

    def _get_priority_vector(self, tensordict) -> torch.Tensor:
        priority = tensordict.get(self.priority_key, None)
        if priority is None:
            return torch.as_tensor(self._sampler.default_priority, dtype=torch.float,
                                   device=tensordict.device).expand(tensordict.shape[0])
        nd = self._storage.ndim
        if nd > 1 and priority.ndim >= nd:
            priority = priority.flatten(0, nd - 1)
        priority = _reduce(priority.reshape(priority.shape[0], -1), self._sampler.reduction, dim=1)
        if nd > 1:
            priority = priority.unflatten(0, tensordict.shape[:nd])
        return priority

    def add(self, data):
        if data is None:
            return torch.zeros((0, self._storage.ndim), dtype=torch.long)
        index = super()._add(data)
        if index is not None:
            if is_tensor_collection(data):
                self._set_index_in_td(data, index)
            self.update_tensordict_priority(data)
        return index

    def extend(self, tensordicts, *, update_priority: bool | None = None) -> torch.Tensor:
        """Extends the replay buffer with a batch of data; when the data holds ``priority_key`` the new items'
        priorities are written right away (``update_priority=False`` disables that)."""
        if not is_tensor_collection(tensordicts):
            raise ValueError(
                f"{self.__class__.__name__} only accepts TensorDictBase subclasses. tensorclasses "
                "and other types are not compatible with that class. Please use a regular `ReplayBuffer` instead.")
        index = super()._extend(tensordicts)
        self._set_index_in_td(tensordicts, index)
        if update_priority is None:
            update_priority = True
        if update_priority:
            try:
                vector = tensordicts.get(self.priority_key, None)
                if vector is not None:
                    self.update_priority(index, vector)
            except Exception as e:
                raise RuntimeError(
                    "Failed to update priority of extended data. You can try to set update_priority=False in the "
                    "extend method and update the priority manually.") from e
        return index

    def _set_index_in_td(self, tensordict, index) -> None:
        if index is None:
            return
        if _is_int(index):
            index = torch.as_tensor(index, device=tensordict.device)
        elif index.ndim == 2 and index.shape[:1] != tensordict.shape[:1]:
            for dim in range(tensordict.ndim, 1, -1):
                if index.shape[:1].numel() == tensordict.shape[:dim].numel():
                    index = index.unflatten(0, tensordict.shape[:dim])
                    break
            else:
                raise RuntimeError(
                    f"could not find how to reshape index with shape {index.shape} to fit in tensordict with "
                    f"shape {tensordict.shape}")
            tensordict.set("index", index)
            return
        tensordict.set("index", expand_as_right(index, tensordict))

    def update_tensordict_priority(self, data) -> None:
        if not isinstance(self._sampler, PrioritizedSampler):
            return
        if data.ndim:
            priority = self._get_priority_vector(data)
        else:
            priority = torch.as_tensor(self._get_priority_item(data))
        index = data.get("index")
        if self._storage.ndim > 1 and index.ndim == 2:
            index = index.unbind(-1)
        else:
            while index.shape != priority.shape:
                index = index[..., 0]
        return self.update_priority(index, priority)

    def sample(self, batch_size: int | None = None, return_info: bool = False, include_info: bool | None = None):
        """Samples a batch; ``"index"`` and the sampler info are written into the returned tensordict."""
        if include_info is not None:
            warnings.warn("include_info is going to be deprecated soon. The default behavior has changed to "
                          "`include_info=True` to avoid bugs linked to wrongly preassigned values in the output "
                          "tensordict.")
        data, info = super().sample(batch_size, return_info=True)
        if is_tensor_collection(data) and include_info in (True, None):
            for key, val in info.items():
                if key == "index" and isinstance(val, tuple):
                    val = torch.stack(val, -1)
                try:
                    val = _to_torch(val, data.device)
                    if val.ndim < data.ndim:
                        val = expand_as_right(val, data)
                    data.set(key, val)
                except RuntimeError:
                    raise RuntimeError(
                        "Failed to set the metadata (e.g., indices or weights) in the sampled tensordict within "
                        "TensorDictReplayBuffer.sample. This is probably caused by a shape mismatch. You can always "
                        "recover these items from the `sample` method from a regular ReplayBuffer instance with "
                        "the 'return_info' flag set to True.")
        elif not is_tensor_collection(data) and include_info in (True, None):
            raise RuntimeError("Cannot include info in non-tensordict data")
        if return_info:
            return data, info
        return data


class TensorDictPrioritizedReplayBuffer(TensorDictReplayBuffer):
    """TensorDict-specific wrapper around :class:`PrioritizedReplayBuffer` (replay_buffers.py:2025-2230).

    The data's ``priority_key`` entry (default ``"td_error"``) drives the priorities:
    ``rb.update_tensordict_priority(sample)`` writes the TD errors of a sampled batch back to the trees.
    """

    def __init__(self, *, alpha: float, beta: float, priority_key: str = "td_error", eps: float = 1e-8,
                 storage=None, sampler_device=None, sync: bool = True, collate_fn=None, reduction: str = "max",
                 batch_size: int | None = None, dim_extend: int | None = None, generator=None, **kwargs) -> None:
        if storage is None:
            storage = ListStorage(max_size=1_000)
        elif not isinstance(storage, Storage) and callable(storage):
            storage = storage()
        if not sync:
            raise NotImplementedError("sync=False (multi-process writers) is outside the B200 hot path")
        sampler = PrioritizedSampler(storage.max_size, alpha, beta, eps, reduction=reduction, device=sampler_device)
        super().__init__(priority_key=priority_key, storage=storage, sampler=sampler, collate_fn=collate_fn,
                         batch_size=batch_size, dim_extend=dim_extend, generator=generator, **kwargs)
