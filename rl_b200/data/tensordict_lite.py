"""A minimal stand-in for ``tensordict.TensorDict`` (the ``tensordict`` package is not vendored by the
reference -- pyproject.toml:46 -- and is absent from this image).

Only what the replay / advantage hot path touches is provided: nested string keys, a leading
``batch_size`` shared by all leaves, ``get`` / ``set`` with tuple keys, indexing along the batch
dimensions, ``to(device)``, ``keys(include_nested, leaves_only)`` and a couple of shape helpers.  When
the real ``tensordict`` is importable the buffers accept its objects as well (duck typing on this
same surface), so user code written against TorchRL keeps working.
"""
from __future__ import annotations

from typing import Any, Iterator

import torch

NestedKey = Any  # str | tuple[str, ...]


def _norm_key(key: NestedKey) -> tuple:
    if type(key) is str:
        return (key,)
    if type(key) is tuple and key and all(type(k) is str for k in key):
        return key
    if isinstance(key, str):
        return (key,)
    if isinstance(key, tuple) and all(isinstance(k, str) for k in key) and key:
        return key
    raise KeyError(f"keys must be strings or non-empty tuples of strings, got {key!r}")


class TensorDict:
    """``TensorDict(source, batch_size, device=None)`` -- nested dict of tensors with a common batch shape."""

    def __init__(self, source: dict | None = None, batch_size=None, device=None):
        self._data: dict[str, Any] = {}
        self._device = torch.device(device) if device is not None else None
        if batch_size is None:
            batch_size = []
        if isinstance(batch_size, int):
            batch_size = [batch_size]
        self._batch_size = torch.Size(batch_size)
        for k, v in (source or {}).items():
            self.set(k, v)

    @classmethod
    def _from_leaves(cls, keys, leaves, batch_size) -> "TensorDict":
        """Trusted fast constructor: nested keys (str | tuple) + leaves that already share `batch_size`."""
        out = cls.__new__(cls)
        out._data = {}
        out._device = None
        out._batch_size = batch_size if isinstance(batch_size, torch.Size) else torch.Size(batch_size)
        for k, v in zip(keys, leaves):
            if isinstance(k, str):
                out._data[k] = v
                continue
            node = out
            for part in k[:-1]:
                nxt = node._data.get(part)
                if nxt is None:
                    nxt = cls.__new__(cls)
                    nxt._data, nxt._device, nxt._batch_size = {}, None, out._batch_size
                    node._data[part] = nxt
                node = nxt
            node._data[k[-1]] = v
        return out

    # ---- metadata
    @property
    def batch_size(self) -> torch.Size:
        return self._batch_size

    @property
    def shape(self) -> torch.Size:
        return self._batch_size

    @property
    def batch_dims(self) -> int:
        return len(self._batch_size)

    @property
    def ndim(self) -> int:
        return len(self._batch_size)

    def __len__(self) -> int:
        if not self._batch_size:
            raise TypeError("len() of a 0-d TensorDict")
        return self._batch_size[0]

    def numel(self) -> int:
        return self._batch_size.numel()

    @property
    def device(self):
        if self._device is not None:
            return self._device
        for v in self._data.values():       # the first leaf decides; no key list is built
            d = v.device
            if d is not None:
                return d
        return None

    @property
    def is_locked(self) -> bool:
        return False

    def unlock_(self):
        return self

    def lock_(self):
        return self

    # ---- access
    def _wrap(self, value):
        if isinstance(value, TensorDict):
            return value
        if isinstance(value, dict):
            return TensorDict(value, self._batch_size, self._device)
        if not isinstance(value, torch.Tensor):
            value = torch.as_tensor(value, device=self._device)
        elif self._device is not None and value.device != self._device:
            value = value.to(self._device)
        return value

    def set(self, key: NestedKey, value) -> "TensorDict":
        key = _norm_key(key)
        node = self
        for k in key[:-1]:
            nxt = node._data.get(k)
            if not isinstance(nxt, TensorDict):
                nxt = TensorDict({}, node._batch_size, node._device)
                node._data[k] = nxt
            node = nxt
        value = node._wrap(value)
        if tuple(value.shape[: len(node._batch_size)]) != tuple(node._batch_size):
            raise RuntimeError(
                f"batch dimension mismatch, got self.batch_size={node._batch_size} and value.shape={value.shape}.")
        node._data[key[-1]] = value
        return self


    def get(self, key: NestedKey, default=...):
        if type(key) is str:                # flat key: one dict probe
            v = self._data.get(key)
            if v is not None:
                return v
        node: Any = self
        for k in _norm_key(key):
            if not isinstance(node, TensorDict) or k not in node._data:
                if default is ...:
                    raise KeyError(f"key {key!r} not found in TensorDict with keys {sorted(self._data)}")
                return default
            node = node._data[k]
        return node

    def pop(self, key: NestedKey, default=...):
        key = _norm_key(key)
        node = self.get(key[:-1]) if len(key) > 1 else self
        if key[-1] in node._data:
            return node._data.pop(key[-1])
        if default is ...:
            raise KeyError(key)
        return default

    def __contains__(self, key) -> bool:
        return self.get(key, None) is not None

    def keys(self, include_nested: bool = False, leaves_only: bool = False) -> list:
        out = []
        for k, v in self._data.items():
            if isinstance(v, TensorDict):
                if not leaves_only:
                    out.append(k)
                if include_nested:
                    out.extend((k, *(_norm_key(s))) for s in v.keys(True, leaves_only))
            else:
                out.append(k)
        return out

    def items(self, include_nested: bool = False, leaves_only: bool = False) -> Iterator:
        for k, v in self._data.items():
            if isinstance(v, TensorDict):
                if not leaves_only:
                    yield k, v
                if include_nested:
                    for sk, sv in v.items(True, leaves_only):
                        yield (k, *((sk,) if type(sk) is str else sk)), sv
            else:
                yield k, v

    def values(self, include_nested: bool = False, leaves_only: bool = False) -> Iterator:
        for _, v in self.items(include_nested, leaves_only):
            yield v

    # ---- transformations
    def apply(self, fn, batch_size=None) -> "TensorDict":
        out = TensorDict({}, self._batch_size if batch_size is None else batch_size, None)
        for k, v in self._data.items():
            out._data[k] = v.apply(fn, batch_size=batch_size) if isinstance(v, TensorDict) else fn(v)
        return out

    def __getitem__(self, index):
        if isinstance(index, str) or (isinstance(index, tuple) and index and all(isinstance(i, str) for i in index)):
            return self.get(index)
        probe = torch.empty(self._batch_size, device="meta")[index]
        return self.apply(lambda t: t[index], batch_size=probe.shape)

    def __setitem__(self, index, value) -> None:
        if isinstance(index, str) or (isinstance(index, tuple) and index and all(isinstance(i, str) for i in index)):
            self.set(index, value)
            return
        if isinstance(value, dict):
            value = TensorDict(value, torch.empty(self._batch_size, device="meta")[index].shape)
        for k, v in value.items(True, True):
            self.get(k)[index] = v

    def to(self, device) -> "TensorDict":
        out = self.apply(lambda t: t.to(device))
        out._device = torch.device(device)
        return out

    def clone(self) -> "TensorDict":
        out = self.apply(lambda t: t.clone())
        out._device = self._device
        return out

    def contiguous(self) -> "TensorDict":
        return self.apply(lambda t: t.contiguous())

    def expand(self, *shape) -> "TensorDict":
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        nb = len(self._batch_size)
        lead = len(shape) - nb
        return self.apply(lambda t: t.expand(*shape, *t.shape[nb:]) if lead >= 0 else t, batch_size=shape)

    def reshape(self, *shape) -> "TensorDict":
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        nb = len(self._batch_size)
        new = torch.empty(self._batch_size, device="meta").reshape(*shape).shape
        return self.apply(lambda t: t.reshape(*new, *t.shape[nb:]), batch_size=new)

    def flatten(self, start: int, end: int) -> "TensorDict":
        new = torch.empty(self._batch_size, device="meta").flatten(start, end).shape
        return self.apply(lambda t: t.flatten(start, end), batch_size=new)

    def unsqueeze(self, dim: int) -> "TensorDict":
        nb = len(self._batch_size)
        if dim < 0:
            dim = nb + 1 + dim
        new = torch.empty(self._batch_size, device="meta").unsqueeze(dim).shape
        return self.apply(lambda t: t.unsqueeze(dim), batch_size=new)

    def select(self, *keys, strict: bool = True) -> "TensorDict":
        out = TensorDict({}, self._batch_size, self._device)
        for k in keys:
            v = self.get(k, None)
            if v is None:
                if strict:
                    raise KeyError(k)
                continue
            out.set(k, v)
        return out

    def to_dict(self) -> dict:
        return {k: (v.to_dict() if isinstance(v, TensorDict) else v) for k, v in self._data.items()}

    def __repr__(self) -> str:
        fields = ", ".join(
            f"{k}: {v!r}" if isinstance(v, TensorDict) else f"{k}: Tensor({tuple(v.shape)}, {v.dtype})"
            for k, v in self._data.items())
        return f"TensorDict({{{fields}}}, batch_size={list(self._batch_size)}, device={self.device})"


def is_tensor_collection(obj) -> bool:
    """True for our TensorDict and for anything quacking like ``tensordict.TensorDictBase``."""
    if isinstance(obj, TensorDict):
        return True
    return hasattr(obj, "batch_size") and hasattr(obj, "keys") and hasattr(obj, "get") and hasattr(obj, "set") \
        and not isinstance(obj, torch.Tensor)


def expand_as_right(t: torch.Tensor, dest) -> torch.Tensor:
    """Append singleton dims on the right of ``t`` then expand to ``dest``'s (batch) shape."""
    shape = dest.shape
    while t.ndim < len(shape):
        t = t.unsqueeze(-1)
    return t.expand(shape)


def stack_tds(items: list, dim: int = 0) -> "TensorDict":
    """torch.stack for a list of TensorDicts with identical keys."""
    first = items[0]
    bs = list(first.batch_size)
    bs.insert(dim, len(items))
    out = TensorDict({}, bs)
    for k in first.keys(True, True):
        out.set(k, torch.stack([it.get(k) for it in items], dim))
    return out
