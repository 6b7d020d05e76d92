"""Replay-buffer storages: the host-side mirror of ``torchrl.data.replay_buffers.storages``.

Same class names, constructor arguments, method names and error behaviour as the reference for the
part of the interface the hot path uses:

    Storage            abstract contract                      storages.py:171-359
    ListStorage        python list, config C1 plumbing         storages.py:361-520
    TensorStorage      pre-allocated [N, ...] leaves           storages.py:522-1263
    LazyTensorStorage  allocated from the first batch          storages.py:1275-1559

What differs is where the bytes move: a tensor-index ``get`` is ONE ``rlb_gather`` launch over every
leaf (csrc/gather.cu) instead of one ``aten::index`` per leaf, and a tensor-cursor ``set`` is one
``rlb_scatter`` launch.  Integer and slice indices are views / contiguous copies and stay torch ops.
Tensor storages are HBM-resident: there is no host-memory gather path.
"""
from __future__ import annotations

import abc
from copy import copy
from typing import Any, Sequence

import numpy as np
import torch
from torch.utils import _pytree as pytree

from .. import ops
from .tensordict_lite import TensorDict, is_tensor_collection
from .utils import INT_CLASSES, _is_int


class DeferredStatus:
    """A device status word the kernels OR error bits into, mirrored to pinned host memory ASYNCHRONOUSLY: ``poll`` never
    synchronises -- it reports bits whose copy has already completed -- so an error raised by the reference immediately
    (IndexError, "non-positive p_sum", ...) surfaces here on a later call of the same object.  ``check`` synchronises."""

    def __init__(self, device):
        self.word = torch.zeros(1, dtype=torch.int32, device=device)
        self._cuda = self.word.is_cuda
        self._host = torch.zeros(1, dtype=torch.int32).pin_memory() if self._cuda else self.word
        self._evt = None

    def arm(self) -> None:
        """Start mirroring the current bits (no-op under CUDA-graph capture and on CPU)."""
        if not self._cuda or torch.cuda.is_current_stream_capturing():
            return
        self._host.copy_(self.word, non_blocking=True)
        if self._evt is None:
            self._evt = torch.cuda.Event()
        self._evt.record()

    def poll(self) -> int:
        if not self._cuda:
            bits = int(self.word)
            self.word.zero_()
            return bits
        if self._evt is None or not self._evt.query():
            return 0
        bits = int(self._host)
        if bits:
            self._host.zero_()
            self.word.zero_()
        return bits

    def check(self) -> int:
        bits = int(self.word.item())
        self.word.zero_()
        if self._cuda:
            self._host.zero_()
        return bits


# ----------------------------------------------------------------------------------------------------
# leaf flattening for Tensor | TensorDict-like | pytree (dict / tuple / list of tensors)
# ----------------------------------------------------------------------------------------------------
def flatten_data(data) -> tuple[list, tuple]:
    if isinstance(data, torch.Tensor):
        return [data], ("tensor",)
    if is_tensor_collection(data):
        keys = list(data.keys(True, True))
        return [data.get(k) for k in keys], ("td", keys, type(data))
    leaves, spec = pytree.tree_flatten(data)
    return leaves, ("pytree", spec)


def unflatten_data(leaves: Sequence[torch.Tensor], spec: tuple, batch_size) -> Any:
    kind = spec[0]
    if kind == "tensor":
        return leaves[0]
    if kind == "td":
        cls = spec[2]
        if cls is TensorDict:
            return TensorDict._from_leaves(spec[1], leaves, batch_size)
        try:
            out = cls({}, batch_size=list(batch_size))
        except Exception:  # a foreign TensorDictBase subclass we cannot construct
            out = TensorDict({}, batch_size)
        for k, v in zip(spec[1], leaves):
            out.set(k, v)
        return out
    return pytree.tree_unflatten(list(leaves), spec[1])


def _batch_len(data) -> int:
    if isinstance(data, torch.Tensor) or is_tensor_collection(data):
        return data.shape[0]
    if isinstance(data, list):
        return len(data)
    return pytree.tree_leaves(data)[0].shape[0]


# ----------------------------------------------------------------------------------------------------
class Storage(abc.ABC):
    """Container of a replay buffer (reference storages.py:171-359).

    Every storage implements ``set``, ``get`` and ``__len__``; samplers only need ``len()``, ``ndim``,
    ``shape`` and ``device`` from it.
    """

    ndim = 1
    max_size: int
    _rng: torch.Generator | None = None

    def __init__(self, max_size: int, checkpointer=None, compilable: bool = False) -> None:
        self.max_size = int(max_size)
        self.checkpointer = checkpointer
        self._compilable = compilable
        self._attached_entities_list: list = []

    # buffers reading from this storage register themselves so that writers can notify them
    @property
    def _attached_entities(self) -> list:
        lst = getattr(self, "_attached_entities_list", None)
        if lst is None:
            lst = self._attached_entities_list = []
        return lst

    def _attached_entities_iter(self):
        return self._attached_entities

    def attach(self, buffer: Any) -> None:
        if buffer not in self._attached_entities:
            self._attached_entities.append(buffer)

    @property
    def _is_full(self) -> bool:
        return len(self) == self.max_size

    @abc.abstractmethod
    def set(self, cursor, data: Any, *, set_cursor: bool = True):
        ...

    @abc.abstractmethod
    def get(self, index) -> Any:
        ...

    @abc.abstractmethod
    def __len__(self) -> int:
        ...

    @abc.abstractmethod
    def state_dict(self) -> dict:
        ...

    @abc.abstractmethod
    def load_state_dict(self, state_dict: dict) -> None:
        ...

    @abc.abstractmethod
    def _empty(self) -> None:
        ...

    @abc.abstractmethod
    def contains(self, item) -> bool:
        ...

    def __getitem__(self, item):
        return self.get(item)

    def __setitem__(self, index, value):
        """Writes without moving the cursor or the length (storages.py:272-274)."""
        return self.set(index, value, set_cursor=False)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __contains__(self, item) -> bool:
        return self.contains(item)

    def _rand_given_ndim(self, batch_size: int):
        # uniform indices for RandomSampler (storages.py:296-306)
        if self.ndim == 1:
            return torch.randint(0, len(self), (batch_size,), generator=self._rng,
                                 device=getattr(self, "device", None))
        raise RuntimeError(
            f"Random number generation is not implemented for storage of type {type(self)} with ndim {self.ndim}.")

    @property
    def shape(self):
        if self.ndim == 1:
            return torch.Size([self.max_size])
        raise RuntimeError(f"storage.shape is not supported for storages of type {type(self)} when ndim > 1.")

    def _max_size_along_dim0(self, *, single_data=None, batched_data=None) -> int:
        if self.ndim == 1:
            return self.max_size
        raise RuntimeError(
            f"storage._max_size_along_dim0 is not supported for storages of type {type(self)} when ndim > 1.")

    def flatten(self):
        if self.ndim == 1:
            return self
        raise RuntimeError(f"storage.flatten is not supported for storages of type {type(self)} when ndim > 1.")

    def dumps(self, path) -> None:
        # storages.py:305-311: the storage's checkpointer decides the on-disk layout (ListStorage: one pickled file)
        if self.checkpointer is not None:
            return self.checkpointer.dumps(self, path)
        import os

        os.makedirs(path, exist_ok=True)
        torch.save(self.state_dict(), os.path.join(str(path), "storage.pt"))

    def loads(self, path) -> None:
        if self.checkpointer is not None:
            return self.checkpointer.loads(self, path)
        import os

        self.load_state_dict(torch.load(os.path.join(str(path), "storage.pt"), weights_only=False))

    save = dump = dumps
    load = loads

    def __getstate__(self):
        state = copy(self.__dict__)
        state["_rng"] = None
        return state


# ----------------------------------------------------------------------------------------------------
class ListStorage(Storage):
    """A storage kept in a python list (reference storages.py:361-520); items are arbitrary objects."""

    def __init__(self, max_size: int | None = None, *, compilable: bool = False, device=None):
        if max_size is None:
            max_size = torch.iinfo(torch.int64).max
        super().__init__(max_size, compilable=compilable)
        self._storage: list = []
        self.device = device

    def _to_device(self, data):
        if self.device is None:
            return data
        if hasattr(data, "to"):
            return data.to(self.device)
        return pytree.tree_map(lambda x: x.to(self.device) if hasattr(x, "to") else x, data)

    def set(self, cursor, data: Any, *, set_cursor: bool = True):
        if isinstance(cursor, INT_CLASSES):
            if cursor > len(self._storage):
                raise RuntimeError(
                    "Cannot append data located more than one item away from the storage size: the storage size "
                    f"is {len(self._storage)} and the index of the item to be set is {cursor}.")
            if cursor >= self.max_size:
                raise RuntimeError(
                    f"Cannot append data to the list storage: maximum capacity is {self.max_size} and the index "
                    f"of the item to be set is {cursor}.")
            data = self._to_device(data)
            if cursor == len(self._storage):
                self._storage.append(data)
            else:
                self._storage[cursor] = data
            return
        if isinstance(cursor, (torch.Tensor, np.ndarray)) and cursor.ndim == 0:
            return self.set(int(cursor), data, set_cursor=set_cursor)
        if isinstance(cursor, slice):
            self._storage[cursor] = self._to_device(data)
            return
        if not (isinstance(data, (list, tuple, torch.Tensor, range, set, np.ndarray)) or is_tensor_collection(data)):
            raise TypeError(
                f"Cannot extend a {type(self)} with data of type {type(data)}. Provide a list, tuple, set, range, "
                "np.ndarray, tensor or tensordict subclass instead.")
        cursor = list(cursor.tolist() if hasattr(cursor, "tolist") else cursor)
        items = [data[i] for i in range(len(data))] if not isinstance(data, (list, tuple)) else list(data)
        if len(cursor) != len(items):
            raise ValueError("cursor and data must have the same length")
        for c, d in zip(cursor, items):
            self.set(int(c), d, set_cursor=set_cursor)

    def get(self, index):
        if isinstance(index, INT_CLASSES):
            return self._storage[index]
        if isinstance(index, slice):
            return self._storage[index]
        if isinstance(index, tuple):
            if len(index) > 1:
                raise RuntimeError(f"{type(self).__name__} can only be indexed with one-length tuples.")
            return self.get(index[0])
        if isinstance(index, torch.Tensor):
            index = index.cpu().tolist()
        if isinstance(index, INT_CLASSES):
            return self._storage[index]
        return [self._storage[i] for i in index]

    def __len__(self) -> int:
        return len(self._storage)

    def state_dict(self) -> dict:
        return {"_storage": [e if not hasattr(e, "state_dict") else e.state_dict() for e in self._storage]}

    def load_state_dict(self, state_dict: dict) -> None:
        self._storage = list(state_dict["_storage"])

    def _empty(self) -> None:
        self._storage = []

    def contains(self, item) -> bool:
        if isinstance(item, INT_CLASSES):
            return 0 <= item < len(self._storage) if item >= 0 else False
        if isinstance(item, torch.Tensor):
            return pytree.tree_map(self.contains, item.tolist())
        raise NotImplementedError(f"type {type(item)} is not supported yet.")

    def __repr__(self) -> str:
        return f"ListStorage(items={self._storage[:5]}{'...' if len(self._storage) > 5 else ''})"


# ----------------------------------------------------------------------------------------------------
class TensorStorage(Storage):
    """Pre-allocated, HBM-resident storage of tensors / tensordicts / pytrees (storages.py:522-1263).

    Args:
        storage: a tensor, TensorDict or pytree whose leaves are ``[max_size, ...]`` tensors, or ``None``
            (then ``max_size`` is required and the leaves are allocated on the first write).
        max_size: number of items along the leading ``ndim`` dimensions.
        device: where the leaves live ("auto": taken from the first data written).
        ndim: how many leading dimensions index items (trajectory storages use 2).
    """

    _storage = None

    def __init__(self, storage, max_size=None, *, device="cpu", ndim: int = 1, compilable: bool = False):
        if not ((storage is None) ^ (max_size is None)):
            if storage is None:
                raise ValueError("Expected storage to be non-null.")
            if max_size != _batch_len(storage):
                raise ValueError(
                    "The max-size and the storage shape mismatch: got "
                    f"max_size={max_size} for a storage of shape {flatten_data(storage)[0][0].shape}.")
        elif storage is not None:
            max_size = _batch_len(storage)
        self.ndim = ndim
        from .checkpointers import TensorStorageCheckpointer  # the reference's default (storages.py:612-613)

        super().__init__(max_size, checkpointer=TensorStorageCheckpointer(), compilable=compilable)
        self.initialized = storage is not None
        self._len = max_size if self.initialized else 0
        if device == "auto":
            self.device = flatten_data(storage)[0][0].device if storage is not None else "auto"
        else:
            self.device = torch.device(device)
            if self.device.type == "cuda" and self.device.index is None:
                self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self._storage = storage
        self._leaves: list | None = None
        self._spec = None
        self._plan = None
        self._last_cursor = None
        self._status = None
        if storage is not None:
            self._bind(storage)

    # ---- layout ------------------------------------------------------------------------------------
    def _bind(self, storage) -> None:
        self._leaves, self._spec = flatten_data(storage)
        self._plan = None
        self._total_shape_value = torch.Size(self._leaves[0].shape[: self.ndim])
        if self.ndim > 1:
            self.max_size = self._total_shape_value.numel()

    @property
    def _total_shape(self):
        return getattr(self, "_total_shape_value", None)

    @property
    def _len_along_dim0(self):
        n = len(self)
        if self.ndim > 1:
            ts = self._total_shape
            if ts is None:
                return None
            n = -(n // -ts[1:].numel())
        return n

    def _max_size_along_dim0(self, *, single_data=None, batched_data=None) -> int:
        if self.ndim == 1:
            return self.max_size
        ts = self._total_shape
        if ts is not None:
            return -(self.max_size // -ts[1:].numel())
        data = single_data if single_data is not None else batched_data
        if data is None:
            raise ValueError("single_data or batched_data must be passed.")
        shape = flatten_data(data)[0][0].shape[: self.ndim] if not is_tensor_collection(data) \
            else data.shape[: self.ndim]
        if batched_data is not None:
            shape = shape[1:]
        return -(self.max_size // -torch.Size(shape).numel())

    @property
    def shape(self):
        ts = self._total_shape
        if ts is None:
            return None
        if self._is_full:
            return ts
        return torch.Size([self._len_along_dim0, *ts[1:]])

    def _rand_given_ndim(self, batch_size: int):
        if self.ndim == 1:
            return super()._rand_given_ndim(batch_size)
        return tuple(torch.randint(d, (batch_size,), generator=self._rng, device=self.device) for d in self.shape)

    def flatten(self):
        if self.ndim == 1:
            return self
        if not self.initialized:
            raise RuntimeError("Cannot flatten a non-initialized storage.")
        n = self._len_along_dim0
        flat = [leaf[:n].flatten(0, self.ndim - 1) for leaf in self._leaves]
        return TensorStorage(unflatten_data(flat, self._spec, flat[0].shape[:1]), device=self.device)

    # ---- writes ------------------------------------------------------------------------------------
    def _init(self, data) -> None:
        raise RuntimeError(f"Cannot write to a {type(self).__name__} that was built without a storage.")

    def _get_new_len(self, data, cursor) -> None:
        nd = self.ndim - int(_is_int(cursor))
        lead = data.shape[:nd] if (isinstance(data, torch.Tensor) or is_tensor_collection(data)) \
            else flatten_data(data)[0][0].shape[:nd]
        self._len = min(self._len + torch.Size(lead).numel(), self.max_size)

    def _cast(self, datum: torch.Tensor, store: torch.Tensor) -> torch.Tensor:
        if datum.device != store.device or datum.dtype != store.dtype:
            datum = datum.to(device=store.device, dtype=store.dtype, non_blocking=True)
        return datum

    def _leaves_for_write(self, cursor, data, set_cursor: bool) -> list:
        """Bookkeeping shared by every write: stacks lists, moves the fill level, allocates on first use and
        returns the data leaves in storage order."""
        if set_cursor:
            self._last_cursor = cursor
        if isinstance(data, list):
            # a list is a stack of per-item elements (replay_buffers.py extend docstring): stack it leaf-wise
            try:
                if is_tensor_collection(data[0]):
                    from .tensordict_lite import stack_tds

                    data = stack_tds(data)
                else:
                    data = pytree.tree_map(lambda *xs: torch.stack(xs), *data)
            except Exception as err:
                raise RuntimeError(
                    "Stacking the elements of the list resulted in an error. "
                    f"Storages of type {type(self)} expect all elements of the list "
                    "to have the same tree structure.") from err
        if set_cursor:
            self._get_new_len(data, cursor)
        if not self.initialized:
            if _is_int(cursor):
                self._init(data)
            elif is_tensor_collection(data):
                self._init(data[0])
            else:
                self._init(pytree.tree_map(lambda x: x[0], data))
        if is_tensor_collection(data) and self._spec[0] == "td":
            # read by the storage's own keys: extra keys (e.g. the writer's "index") that the storage does not
            # hold are dropped, missing keys are an error -- the reference's locked-storage behaviour
            # (storages.py:1070-1072)
            leaves = []
            for k in self._spec[1]:
                v = data.get(k, None)
                if v is None:
                    raise KeyError(f"key {k} of the storage is missing from the data written to it")
                leaves.append(v)
            return leaves
        leaves, _ = flatten_data(data)
        if len(leaves) != len(self._leaves):
            raise RuntimeError("the data written to the storage does not match its tree structure")
        return leaves

    def _fits_range(self, n: int, data) -> bool:
        """Whether a writer batch of ``n`` items can take ``_extend_range``."""
        return self.ndim == 1 and 0 < n <= self.max_size and (
            is_tensor_collection(data) or isinstance(data, (torch.Tensor, dict, tuple, list)))

    def _extend_range(self, cursor: int, n: int, data, trees=None) -> None:
        """The writer's batch -- rows (cursor + arange(n)) % max_size -- in ONE launch, together with the sampler's
        default-priority write when ``trees`` (an ``ops.RangeUpdate``) is given: ``rlb_extend``, the fused form of
        storages.py:1028-1096 + samplers.py:1093-1096."""
        max0 = self.max_size
        index = slice(cursor, cursor + n) if cursor + n <= max0 else (torch.arange(cursor, cursor + n) % max0)
        leaves = self._leaves_for_write(index, data, True)
        stores = self._leaves
        cast = [self._cast(d, s) for d, s in zip(leaves, stores)]
        for d, s in zip(cast, stores):
            if d.shape[0] != n or d.shape[1:] != s.shape[1:]:
                raise RuntimeError(f"cannot write data of shape {tuple(d.shape)} into storage rows {tuple(s.shape[1:])}")
        ops.backend().extend(stores, [d if d.is_contiguous() or d.ndim == 1 or d[0].is_contiguous() else d.contiguous()
                                      for d in cast], cursor, n, max0, trees)

    def set(self, cursor, data, *, set_cursor: bool = True):
        leaves = self._leaves_for_write(cursor, data, set_cursor)
        if _is_int(cursor) or isinstance(cursor, slice):
            for datum, store in zip(leaves, self._leaves):
                store[cursor] = self._cast(datum, store)
            return
        if isinstance(cursor, tuple):
            cursor = self._linear_index(cursor)
            stores = [leaf.flatten(0, self.ndim - 1) for leaf in self._leaves]
            leaves = [d.flatten(0, d.ndim - s.ndim) if d.ndim > s.ndim else d for d, s in zip(leaves, stores)]
        else:
            cursor = torch.as_tensor(cursor, dtype=torch.long)
            stores = self._leaves
            if cursor.ndim > 1:
                raise RuntimeError("tensor cursors must be one-dimensional")
        cursor = cursor.to(stores[0].device)
        st = self._index_status()
        ops.backend().scatter(stores, [self._cast(d, s) for d, s in zip(leaves, stores)], cursor,
                              stores[0].shape[0], status=None if st is None else st.word)
        if st is not None:
            st.arm()

    # ---- reads -------------------------------------------------------------------------------------
    def _linear_index(self, index: tuple) -> torch.Tensor:
        ts = self._total_shape
        if len(index) != self.ndim:
            raise RuntimeError(f"expected a tuple of {self.ndim} index tensors")
        lin = None
        for d, ix in enumerate(index):
            ix = torch.as_tensor(ix, dtype=torch.long, device=self._leaves[0].device)
            lin = ix if lin is None else lin * ts[d] + ix
        return lin

    def _index_status(self):
        """The status word of user-facing tensor indexing (on by default on CUDA storages).  Polling it first raises --
        one call late, without a sync -- the IndexError torch indexing would have raised at once."""
        if self._status is False:
            return None
        if self._status is None:
            dev = self._leaves[0].device
            if dev.type != "cuda":
                self._status = False
                return None
            self._status = DeferredStatus(dev)
        if self._status._cuda and torch.cuda.is_current_stream_capturing():
            return self._status    # (an event query would invalidate a capture)
        if self._status.poll() & ops.STATUS_INDEX_OOB:
            raise IndexError("index out of range in an earlier tensor-indexed read / write of this storage "
                             "(reads returned the clamped row, writes were dropped)")
        return self._status

    def _get_trusted(self, index: torch.Tensor):
        """``get`` for indices a sampler of this engine produced (in range by construction): no status bookkeeping."""
        if self.ndim > 1 or not (isinstance(index, torch.Tensor) and index.ndim == 1 and index.dtype == torch.int64
                                 and index.device == self._leaves[0].device):
            return self.get(index)
        if self._plan is None:
            self._plan = ops.backend().gather_plan(self._leaves)
        return unflatten_data(self._plan.run(index, self._len_along_dim0), self._spec, index.shape)

    def get(self, index):
        if not self.initialized:
            raise RuntimeError("Cannot get elements out of a non-initialized storage.")
        n0 = self._len_along_dim0
        if _is_int(index) or isinstance(index, slice) or index is None or index is Ellipsis:
            # views of the filled part; no bytes move
            out = [leaf[:n0][index] for leaf in self._leaves]
            lead = torch.empty((n0, *self._total_shape[1:]), device="meta")[index].shape
            return unflatten_data(out, self._spec, lead)
        be = ops.backend()
        if isinstance(index, tuple):
            if len(index) == 1:
                return self.get(index[0])
            lin = self._linear_index(index)
            leaves = [leaf.flatten(0, self.ndim - 1) for leaf in self._leaves]
            length = n0 * self._total_shape[1:].numel()
            st = self._index_status()
            out = be.gather(leaves, lin.reshape(-1), length, status=None if st is None else st.word)
            if st is not None:
                st.arm()
            out = [o.reshape(*lin.shape, *o.shape[1:]) for o in out]
            return unflatten_data(out, self._spec, lin.shape)
        if not (isinstance(index, torch.Tensor) and index.dtype == torch.int64 and index.device == self._leaves[0].device):
            index = torch.as_tensor(index)
            if index.dtype == torch.bool:
                index = index.nonzero().squeeze(-1)
            index = index.to(device=self._leaves[0].device, dtype=torch.long)
        if self.ndim > 1:
            # a 1-d tensor index on a multi-dim storage selects whole dim-0 slabs: plain torch indexing
            out = [leaf[:n0][index] for leaf in self._leaves]
            return unflatten_data(out, self._spec, out[0].shape[: index.ndim + self.ndim - 1])
        if self._plan is None:
            self._plan = be.gather_plan(self._leaves)
        st = self._index_status()
        word = None if st is None else st.word
        if index.ndim == 1:
            out = self._plan.run(index, n0, status=word)
        else:
            out = self._plan.run(index.reshape(-1), n0, status=word)
            out = [o.reshape(*index.shape, *o.shape[1:]) for o in out]
        if st is not None:
            st.arm()
        return unflatten_data(out, self._spec, index.shape)

    def enable_index_check(self, enabled: bool = True) -> None:
        """IndexError parity for tensor indices (ON by default on CUDA): kernels record out-of-range indices in a device
        word; the next tensor-indexed call of this storage polls its asynchronous host mirror and raises, and
        ``check_index_status`` synchronises and raises now."""
        self._status = None if enabled else False

    def check_index_status(self) -> None:
        if isinstance(self._status, DeferredStatus) and self._status.check() & ops.STATUS_INDEX_OOB:
            raise IndexError("index out of range in storage gather / scatter")

    def __len__(self) -> int:
        return self._len

    def _empty(self) -> None:
        # the layout stays; only the fill level resets (storages.py:1270-1273)
        self._len = 0

    def contains(self, item) -> bool:
        if isinstance(item, INT_CLASSES):
            return 0 <= item < len(self)
        if isinstance(item, torch.Tensor):
            return (item >= 0) & (item < len(self))
        raise NotImplementedError(f"type {type(item)} is not supported yet.")

    def state_dict(self) -> dict:
        return {"_storage": None if self._leaves is None else [l.detach().cpu() for l in self._leaves],
                "_spec": self._spec, "initialized": self.initialized, "_len": self._len}

    def load_state_dict(self, state_dict: dict) -> None:
        leaves = state_dict["_storage"]
        if leaves is not None:
            if self._leaves is None:
                dev = self.device if self.device != "auto" else leaves[0].device
                self._leaves = [l.to(dev) for l in leaves]
                self._spec = state_dict["_spec"]
                self._storage = unflatten_data(self._leaves, self._spec, self._leaves[0].shape[: self.ndim])
                self._total_shape_value = torch.Size(self._leaves[0].shape[: self.ndim])
            else:
                for dst, src in zip(self._leaves, leaves):
                    dst.copy_(src)
        self.initialized = state_dict["initialized"]
        self._len = state_dict["_len"]

    def __repr__(self) -> str:
        return f"{type(self).__name__}(max_size={self.max_size}, len={len(self)}, device={self.device}, ndim={self.ndim})"


class LazyTensorStorage(TensorStorage):
    """Storage whose leaves are allocated from the first batch written (storages.py:1275-1559).

    Args:
        max_size: capacity (number of items).
        device: where to allocate ("auto": the device of the first data).  HBM-resident by design:
            a 1M-frame Atari buffer is 28.2 GB per pixel leaf and a B200 has 180 GB.
        ndim: number of leading dimensions that index items.
    """

    def __init__(self, max_size: int, *, device="cpu", ndim: int = 1, compilable: bool = False,
                 consolidated: bool = False):
        super().__init__(storage=None, max_size=max_size, device=device, ndim=ndim, compilable=compilable)
        if consolidated:
            raise ValueError("consolidated storages need the tensordict package")

    def _init(self, data) -> None:
        # torch.empty_like(data.expand(max_size, ...)) per leaf (storages.py:1479-1526)
        leaves, spec = flatten_data(data)
        if self.device == "auto":
            self.device = leaves[0].device

        def alloc(x: torch.Tensor) -> torch.Tensor:
            if self.ndim > 1:
                lead = x.shape[: self.ndim - 1]
                n0 = -(self.max_size // -torch.Size(lead).numel())
                return torch.empty((n0, *x.shape), dtype=x.dtype, device=self.device)
            return torch.empty((self.max_size, *x.shape), dtype=x.dtype, device=self.device)

        out = [alloc(x) for x in leaves]
        self._storage = unflatten_data(out, spec, out[0].shape[: self.ndim])
        self._bind(self._storage)
        self.initialized = True
