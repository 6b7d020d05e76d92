"""Samplers: the host-side mirror of ``torchrl.data.replay_buffers.samplers`` for the hot path.

    Sampler             abstract contract                       samplers.py:99-171
    RandomSampler       uniform with replacement                samplers.py:174-218
    SamplerWithoutReplacement  permutation sweeps (PPO epochs)  samplers.py:221-362
    PrioritizedSampler  proportional PER, sum/min segment trees samplers.py:577-1205
    SliceSampler        trajectory slices of a ring of steps    samplers.py:1207-2300 (1-d storages, span=False)

``PrioritizedSampler`` keeps the reference's constructor, properties, bookkeeping quirks (double ``pow``
on default priorities, running max of raw priorities, SURVEY.md 8a') and error messages, but its trees
live in HBM and its arithmetic is two kernel launches:

    sample           torch.rand(B, generator)  ->  rlb_per_sample   (query x2, mass, descent, clamp, leaf,
                                                                     importance weight in ONE launch)
    update_priority  rlb_per_update  ((p+eps)**alpha, running max, last-writer-wins scatter, touched-ancestor
                                      recomputation of both trees)

``torch.rand`` stays a torch call so the random stream is literally the reference's (samplers.py:918).
"""
from __future__ import annotations

import abc
import json
from copy import deepcopy
from pathlib import Path
from typing import Any

import numpy as np
import torch

from .. import ops
from .segment_tree import MinSegmentTreeFp32, MinSegmentTreeFp64, SumSegmentTreeFp32, SumSegmentTreeFp64
from .storages import Storage
from .utils import unravel_index

_EMPTY_STORAGE_ERROR = "Cannot sample from an empty storage."



def _contents(storage):
    """The stored leaves the slice samplers read their signals from (``storage[:]`` in the reference).  Storages whose
    full read would have to rebuild something (``FrameStackStorage``: every frame stack) expose the cheap part as
    ``_signal_view()``."""
    view = getattr(storage, "_signal_view", None)
    return view() if view is not None else storage[:]

class Sampler(abc.ABC):
    """A generic sampler base class for composable replay buffers (samplers.py:99-171)."""

    _rng: torch.Generator | None = None

    @abc.abstractmethod
    def sample(self, storage: Storage, batch_size: int) -> tuple[Any, dict]:
        ...

    def add(self, index: int) -> None:
        return

    def extend(self, index) -> None:
        return

    def update_priority(self, index, priority, *, storage: Storage | None = None) -> dict | None:
        return

    def mark_update(self, index, *, storage: Storage | None = None) -> None:
        return

    @property
    def default_priority(self) -> float:
        return 1.0

    @abc.abstractmethod
    def state_dict(self) -> dict:
        ...

    @abc.abstractmethod
    def load_state_dict(self, state_dict: dict) -> None:
        ...

    @property
    def ran_out(self) -> bool:
        # by default, samplers never run out
        return False

    @abc.abstractmethod
    def _empty(self) -> None:
        ...

    @abc.abstractmethod
    def dumps(self, path) -> None:
        ...

    @abc.abstractmethod
    def loads(self, path) -> None:
        ...

    @property
    def _remaining_batches(self) -> int:
        return torch.iinfo(torch.int64).max

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_rng"] = None
        return state


class RandomSampler(Sampler):
    """A uniformly random sampler with replacement (samplers.py:174-218)."""

    def sample(self, storage: Storage, batch_size: int) -> tuple[torch.Tensor, dict]:
        if len(storage) == 0:
            raise RuntimeError(_EMPTY_STORAGE_ERROR)
        return storage._rand_given_ndim(batch_size), {}

    def _empty(self) -> None:
        pass

    def dumps(self, path) -> None:
        pass

    def loads(self, path) -> None:
        pass

    def state_dict(self) -> dict:
        return {}

    def load_state_dict(self, state_dict: dict) -> None:
        return


class SamplerWithoutReplacement(Sampler):
    """A data-consuming sampler: consecutive batches never share an item until the storage has been swept
    (samplers.py:221-362; PPO minibatching, sota-implementations/ppo/ppo_atari.py:83-91).

    Args:
        drop_last (bool, optional): if ``True`` an incomplete last batch is dropped (``ran_out`` turns true when the
            remaining indices cannot fill a batch). Defaults to ``False``: the last batch of a sweep may be short.
        shuffle (bool, optional): if ``False`` items come in storage order. Defaults to ``True``.

    A fresh permutation (``torch.randperm(len, generator)`` on the storage device -- the reference's RNG call) is drawn
    whenever the storage length changes or the previous one is exhausted.  Index generation is a handful of tiny
    torch ops per sweep; the bytes are moved by the storage's gather kernel.
    """

    def __init__(self, drop_last: bool = False, shuffle: bool = True):
        self._sample_list = None
        self.len_storage = 0
        self.drop_last = drop_last
        self._ran_out = False
        self.shuffle = shuffle
        self._remaining = torch.iinfo(torch.int64).max

    @property
    def _remaining_batches(self) -> int:
        return self._remaining

    def _count_remaining(self, batch_size: int) -> None:
        n = self._sample_list.numel()
        self._remaining = n // batch_size if self.drop_last else -(n // -batch_size)

    def _new_order(self, storage, len_storage: int, batch_size: int) -> None:
        device = self._sample_list.device if storage is None else getattr(storage, "device", None)
        if device == "auto":
            device = None
        if self.shuffle:
            self._sample_list = torch.randperm(len_storage, device=device, generator=self._rng)
        else:
            self._sample_list = torch.arange(len_storage, device=device)
        self._count_remaining(batch_size)

    def _storage_len(self, storage) -> int:
        return len(storage)

    def sample(self, storage: Storage, batch_size: int) -> tuple[Any, dict]:
        len_storage = self._storage_len(storage)
        if len_storage == 0:
            raise RuntimeError(_EMPTY_STORAGE_ERROR)
        if self.len_storage != len_storage or self._sample_list is None:
            self._new_order(storage, len_storage, batch_size)
        if len_storage < batch_size and self.drop_last:
            raise ValueError(
                f"The batch size ({batch_size}) is greater than the storage capacity ({len_storage}). "
                "This makes it impossible to return a sample without repeating indices. "
                "Consider changing the sampler class or turn the 'drop_last' argument to False.")
        self.len_storage = len_storage
        index = self._sample_list[:batch_size]
        self._sample_list = self._sample_list[batch_size:]
        self._count_remaining(batch_size)
        left = self._sample_list.shape[0]
        if left == 0 or (self.drop_last and left < batch_size):
            self.ran_out = True  # read by ReplayBuffer.__iter__ as the end of the sweep
            self._new_order(None, len_storage, batch_size)
        else:
            self.ran_out = False
        if storage.ndim > 1:
            index = unravel_index(index, storage.shape)
        return index, {}

    @property
    def ran_out(self) -> bool:
        return self._ran_out

    @ran_out.setter
    def ran_out(self, value: bool) -> None:
        self._ran_out = value

    def _empty(self) -> None:
        self._sample_list = None
        self.len_storage = 0
        self._ran_out = False

    def state_dict(self) -> dict:
        return {"len_storage": self.len_storage, "_sample_list": self._sample_list, "drop_last": self.drop_last,
                "_ran_out": self._ran_out}

    def load_state_dict(self, state_dict: dict) -> None:
        self.len_storage = state_dict["len_storage"]
        self._sample_list = state_dict["_sample_list"]
        self.drop_last = state_dict["drop_last"]
        self._ran_out = state_dict["_ran_out"]

    def dumps(self, path) -> None:
        path = Path(path)
        path.mkdir(exist_ok=True, parents=True)
        torch.save(self.state_dict(), path / "sampler_state.pt")

    def loads(self, path) -> None:
        self.load_state_dict(torch.load(Path(path) / "sampler_state.pt", weights_only=False))

    def __repr__(self) -> str:
        perc = len(self._sample_list) / self.len_storage * 100 if self._sample_list is not None else 0.0
        return f"{self.__class__.__name__}({perc: 4.4f}% sampled)"


class SliceSampler(Sampler):
    """Samples slices of data along the first dimension, given start and stop signals (samplers.py:1207-2300).

    Keyword Args:
        num_slices (int): the number of slices to be sampled; the batch size must be divisible by it. Exclusive with
            ``slice_len``.
        slice_len (int): the length of the slices to be sampled; the batch size must be divisible by it.
        end_key (NestedKey, optional): the key indicating the end of a trajectory. Defaults to ``("next", "done")``.
        traj_key (NestedKey, optional): the key indicating the trajectories. When neither key is given,
            ``("collector", "traj_ids")`` then ``"episode"`` are looked up in the storage, then ``end_key`` is used.
        ends (torch.Tensor, optional): a 1d boolean tensor with the end-of-trajectory signals (needs ``cache_values``).
        trajectories (torch.Tensor, optional): a 1d integer tensor with the trajectory ids (needs ``cache_values``).
        cache_values (bool, optional): keep the trajectory table until the buffer is written to. Defaults to ``False``.
        truncated_key (NestedKey, optional): if not ``None``, the last step of every slice is marked truncated (and done)
            in the info written into the sample. Defaults to ``("next", "truncated")``.
        strict_length (bool, optional): if ``False``, trajectories shorter than the slice length are sampled whole and
            the batch may be shorter than asked. Defaults to ``True`` (they are never sampled).
        pad_output (bool, optional): with ``strict_length=False``, pad short slices to the slice length by repeating
            their last step and return ``("collector", "mask")``.
        span (bool, int, Tuple[bool | int, bool | int], optional): let a slice hang out of its trajectory on the left
            and / or on the right (``True``: by up to ``slice_len - 1`` steps, an int: by up to that many); the part
            outside is cut off, so such slices are shorter (:2071-2118).  Makes slice lengths variable, like
            ``strict_length=False``.
        compile, use_gpu: accepted for signature compatibility; no effect (everything already runs on the storage's
            device).

    Storages with ``ndim`` 1 or 2 (``[T, E]``: one ring per column, the table is built ring by ring -- the order of the
    reference's transposed ``nonzero`` -- and the sample is a ``(time, column)`` index pair).  Both halves of the
    reference's index arithmetic are kernels: ``rlb_traj_table``
    (trajectory boundaries of the ring; the reference's nonzero / roll / boolean-index sequence, :1652-1743, :1993-2010)
    and ``rlb_slice_index`` (slice expansion, :2058-2215).  The two random draws are the reference's own calls
    (``torch.randint(n_trajectories, (num_slices,))`` then ``torch.rand(num_slices)``), so a seeded generator yields the
    reference's slices.  Like the reference, building the table reads two counters back (one sync); with
    ``cache_values=True`` that happens once per write instead of once per sample.
    """

    def __init__(self, *, num_slices: int | None = None, slice_len: int | None = None, end_key=None, traj_key=None,
                 ends: torch.Tensor | None = None, trajectories: torch.Tensor | None = None,
                 cache_values: bool = False, truncated_key=("next", "truncated"), strict_length: bool = True,
                 pad_output: bool = False, compile=False, span=False, use_gpu=False):
        self.num_slices, self.slice_len = num_slices, slice_len
        self.end_key, self.traj_key = end_key, traj_key
        self.truncated_key = truncated_key
        self.cache_values = cache_values
        self.strict_length = strict_length
        if pad_output and strict_length:
            raise ValueError(
                "pad_output=True is incompatible with strict_length=True: padding only happens when short trajectories "
                "are kept, which requires strict_length=False.")
        self.pad_output = pad_output
        if isinstance(span, (bool, int)):
            span = (span, span)
        self.span = tuple(span)
        # the kernel's encoding: 0 = off, -1 = True, k > 0 = at most k steps outside the trajectory
        self._span_code = tuple(0 if not v else (-1 if v is True else int(v)) for v in self.span)
        self._cache: dict = {}
        self._given = None          # (signal tensor, by_id) passed to the constructor
        self._fetch_traj, self._traj_key_auto = True, False
        if trajectories is not None or ends is not None:
            what = "trajectories" if trajectories is not None else "ends"
            if traj_key is not None or end_key:
                raise RuntimeError(f"`{what}` and `end_key` or `traj_key` are exclusive arguments.")
            if trajectories is not None and ends is not None:
                raise RuntimeError("trajectories and ends are exclusive arguments.")
            if not cache_values:
                raise RuntimeError(f"To be used, {what} requires `cache_values` to be set to `True`.")
            sig = trajectories if trajectories is not None else ends
            if sig.ndim > 2:
                raise NotImplementedError("SliceSampler supports storages with ndim 1 or 2")
            self._given = (sig, trajectories is not None)
        else:
            if traj_key is not None:
                self._fetch_traj = True
            elif end_key is not None:
                self._fetch_traj = False
            else:
                self._traj_key_auto = True
            self.end_key = end_key if end_key is not None else ("next", "done")
        if not ((num_slices is None) ^ (slice_len is None)):
            raise TypeError("Either num_slices or slice_len must be not None, and not both. "
                            f"Got num_slices={num_slices} and slice_len={slice_len}.")
        self._traj_buf = self._traj_counts = self._traj_ws = None
        self._traj_ws_len = 0
        self._column = None          # N-d storages: the ring (column) of every table entry

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(num_slices={self.num_slices}, slice_len={self.slice_len}, "
                f"end_key={self.end_key}, traj_key={self.traj_key}, truncated_key={self.truncated_key}, "
                f"strict_length={self.strict_length}, pad_output={self.pad_output})")

    def extend(self, index) -> None:
        super().extend(index)
        if self._given is None:
            self._cache.clear()

    def add(self, index) -> None:
        super().add(index)
        if self._given is None:
            self._cache.clear()

    def _empty(self) -> None:
        self._cache.clear()

    def dumps(self, path) -> None:   # no-op: the table is derived from the storage
        ...

    def loads(self, path) -> None:
        ...

    def state_dict(self) -> dict:
        return {}

    def load_state_dict(self, state_dict: dict) -> None:
        ...

    # ---- trajectory table ---------------------------------------------------------------------------
    def _resolve_traj_key(self, contents) -> None:
        # samplers.py:1798-1864: prefer what collectors write, then "episode", then reconstruct from end_key
        self._traj_key_auto = False
        keys = set(contents.keys(True, True)) if hasattr(contents, "keys") else set()
        if ("collector", "traj_ids") in keys:
            self.traj_key, self._fetch_traj = ("collector", "traj_ids"), True
        elif "episode" in keys:
            self.traj_key, self._fetch_traj = "episode", True
        else:
            self._fetch_traj = False

    def _signal(self, storage):
        """(signal tensor over the filled slots, by_id, at_capacity, cursor)."""
        if self._given is not None:   # the reference assumes a full storage here (:1565-1586)
            sig, by_id = self._given
            return sig.to(storage.device) if hasattr(storage, "device") else sig, by_id, True, -1
        try:
            contents = _contents(storage)
        except Exception:
            raise RuntimeError("Could not get a tensordict out of the storage, which is required for SliceSampler to "
                               "compute the trajectories.")
        if self._traj_key_auto:
            self._resolve_traj_key(contents)
        sig = None
        for attempt in range(2):
            key = self.traj_key if self._fetch_traj else self.end_key
            try:
                sig = contents.get(key)
                break
            except (KeyError, AttributeError):
                if attempt or (not self._fetch_traj and self.traj_key is None):
                    raise KeyError(f"SliceSampler could not find {key!r} in the storage")
                self._fetch_traj = not self._fetch_traj      # fall back to the other signal (:1893-1927)
        nd = getattr(storage, "ndim", 1)
        while sig.ndim > nd and sig.shape[-1] == 1:
            sig = sig.squeeze(-1)
        if sig.ndim != nd or nd > 2:
            raise NotImplementedError(
                f"SliceSampler on the B200 engine supports storages with ndim 1 or 2; got a signal of shape "
                f"{tuple(sig.shape)} for ndim={nd}")
        cursor = getattr(storage, "_last_cursor", None)
        if isinstance(cursor, slice):
            cursor = cursor.stop - 1
        elif isinstance(cursor, range):
            cursor = cursor[-1]
        elif isinstance(cursor, torch.Tensor):
            cursor = int(cursor.reshape(-1)[-1])
        if isinstance(cursor, tuple):       # N-d writes address (row, ...) tuples: the row is what closes a trajectory
            cursor = cursor[0]
            cursor = int(cursor.reshape(-1)[-1]) if isinstance(cursor, torch.Tensor) else cursor
        n0 = storage.shape[0] if nd > 1 else len(storage)
        return sig[:n0], self._fetch_traj, bool(storage._is_full), -1 if cursor is None else int(cursor)

    def _table_nd(self, sig, by_id, at_capacity, cursor, seq_length: int):
        """[T, E] signal: one ring per column.  Every column goes through ``rlb_traj_table`` on its own (a transposing
        copy makes it contiguous); all counters come back in ONE read, then the per-column tables are concatenated in
        column order -- the order of the reference's ``end.transpose(0, -1).nonzero()`` (:1717-1718).  Returns
        (table [3, n], column int64[n], n_trajectories, n_long_enough)."""
        T, E = sig.shape
        dev = sig.device
        be = ops.backend()
        cols = sig.t().contiguous()
        tables = torch.empty((E, 3, T), dtype=torch.int64, device=dev)
        counts = torch.zeros((E, 2), dtype=torch.int64, device=dev)
        if self._traj_ws is None or self._traj_ws_len < T or self._traj_ws.device != dev:
            self._traj_ws, self._traj_ws_len = be.traj_workspace(T, dev), T
        for e in range(E):
            be.traj_table(cols[e], by_id, T, at_capacity, cursor, seq_length, self.strict_length, tables[e], counts[e],
                          self._traj_ws)
        cnt = counts.tolist()                                                # the one synchronisation of a table build
        keep = [c[1] if self.strict_length else c[0] for c in cnt]
        table = torch.cat([tables[e, :, :k] for e, k in enumerate(keep)], dim=1)
        column = torch.repeat_interleave(torch.arange(E, device=dev), torch.tensor(keep, device=dev))
        return table, column, sum(c[0] for c in cnt), sum(c[1] for c in cnt)

    def _table(self, storage, seq_length: int):
        """(table int64 [3, L] = start / stop / length rows, n_trajectories, n_long_enough); N-d storages also cache the
        column of every trajectory (``self._cache[("column", seq_length)]``)."""
        key = ("table", seq_length)
        if self.cache_values and key in self._cache:
            return self._cache[key]
        sig, by_id, at_capacity, cursor = self._signal(storage)
        L = sig.shape[0]
        if L == 0:
            raise RuntimeError(_EMPTY_STORAGE_ERROR)
        if sig.ndim == 2:
            table, column, n_all, n_long = self._table_nd(sig, by_id, at_capacity, cursor, seq_length)
            self._column = column
            out = (table, n_all, n_long)
            if self.cache_values:
                self._cache[key] = out
            return out
        dev = sig.device
        be = ops.backend()
        if self._traj_buf is None or self._traj_buf.shape[1] < L or self._traj_buf.device != dev:
            size = max(L, getattr(storage, "max_size", L))
            self._traj_buf = torch.empty((3, size), dtype=torch.int64, device=dev)
            self._traj_counts = torch.zeros(2, dtype=torch.int64, device=dev)
            self._traj_ws, self._traj_ws_len = be.traj_workspace(size, dev), size
        table = self._traj_buf if not self.cache_values else torch.empty_like(self._traj_buf)
        be.traj_table(sig, by_id, L, at_capacity, cursor, seq_length, self.strict_length, table, self._traj_counts,
                      self._traj_ws)
        n_all, n_long = (int(c) for c in self._traj_counts.tolist())      # the one synchronisation of a table build
        out = (table, n_all, n_long)
        if self.cache_values:
            self._cache[key] = out
        return out

    def _plan(self, storage, seq_length: int, num_slices: int):
        """(table, n_trajectories a slice may come from, variable-length flag, trajectory ids or None = draw them)."""
        table, n_all, n_long = self._table(storage, seq_length)
        if self.strict_length:
            if n_long == 0:
                raise RuntimeError("Did not find a single trajectory with sufficient length "
                                   f"(required={seq_length}, trajectories={n_all}).")
            return table, n_long, False, None
        return table, n_all, n_long < n_all, None

    def _adjusted_batch_size(self, batch_size: int) -> tuple[int, int]:
        if self.num_slices is not None:
            if batch_size % self.num_slices != 0:
                raise RuntimeError("The batch-size must be divisible by the number of slices, got "
                                   f"batch_size={batch_size} and num_slices={self.num_slices}.")
            return batch_size // self.num_slices, self.num_slices
        if batch_size % self.slice_len != 0:
            raise RuntimeError("The batch-size must be divisible by the slice length, got "
                               f"batch_size={batch_size} and slice_len={self.slice_len}.")
        return self.slice_len, batch_size // self.slice_len

    # ---- sample (samplers.py:1947-2215) -------------------------------------------------------------
    def sample(self, storage: Storage, batch_size: int) -> tuple[Any, dict]:
        nd = storage.ndim
        if nd > 2:
            raise NotImplementedError("SliceSampler on the B200 engine supports storages with ndim 1 or 2")
        seq_length, num_slices = self._adjusted_batch_size(batch_size)
        span = self._span_code
        if any(k >= seq_length for k in span):
            raise ValueError("The right and left span must be strictly lower than the sequence length")
        table, n_traj, variable, traj = self._plan(storage, seq_length, num_slices)
        variable = variable or any(span)        # a span cuts slices short where they leave their trajectory
        dev = table.device
        if traj is None:
            traj = torch.randint(n_traj, (num_slices,), device=dev, generator=self._rng)    # :1987-1990
        num_slices = traj.numel()
        u = torch.rand(num_slices, device=dev, generator=self._rng)                          # :2099-2102
        be = ops.backend()
        storage_length = storage.shape[0]
        args = (table[0], table[2], n_traj, traj, u, seq_length, storage_length)
        # the stored done / terminated flags of the sampled steps ride in the same launch when they are one byte per slot
        contents = _contents(storage) if hasattr(storage, "get") else None
        get = (lambda k: contents.get(k, None)) if contents is not None and hasattr(contents, "get") else (lambda k: None)
        done_all = term_all = None
        if self.truncated_key is not None:
            done_key = _replace_last(self.truncated_key, "done")
            terminated_key = _replace_last(self.truncated_key, "terminated")
            done_all, term_all = get(done_key), get(terminated_key)
        one_byte = lambda t: t is None or (t.element_size() == 1 and t.numel() == t.shape[0] and t.is_contiguous())
        fused = nd == 1 and self.truncated_key is not None and one_byte(done_all) and one_byte(term_all)
        kw = dict(flags=(done_all, term_all)) if fused else {}
        if variable and not self.pad_output:
            seq = be.slice_index(*args, variable=True, want_index=False, span=span)[3]
            ends_at = seq.cumsum(0)
            total = int(ends_at[-1])                                                         # data-dependent batch size
            out = be.slice_index(*args, variable=True, out_offset=ends_at - seq, total=total, span=span, **kw)
            slice_starts = ends_at - seq
            per_slice = seq
        else:
            out = be.slice_index(*args, variable=variable, pad_output=self.pad_output, span=span, **kw)
            slice_starts = None
            per_slice = None
        index, truncated, mask = out[0], out[1], out[2]
        flat = index
        if nd == 2:
            # the column of every sampled step: that of its trajectory; rows of the [T, E, ...] leaves are t * E + e
            col = self._column[traj]
            col = col.repeat_interleave(per_slice) if per_slice is not None else col.repeat_interleave(seq_length)
            flat = index * storage.shape[1] + col
        info: dict = {}
        if mask is not None:
            info[("collector", "mask")] = mask
        is_init_all = get("is_init")
        n_rows = len(storage)
        rows_of = (lambda t: t.reshape(-1, *t.shape[nd:])) if nd > 1 else (lambda t: t)
        if self.truncated_key is not None:
            info[self.truncated_key] = truncated
            if fused:
                info[done_key], info[terminated_key] = out[4], out[5]
            else:     # wide flag leaves: one small gather, then the reference's elementwise ops (:2190-2205)
                have = {k: rows_of(v) for k, v in (("done", done_all), ("terminated", term_all)) if v is not None}
                rows = dict(zip(have, be.gather(list(have.values()), flat, n_rows))) if have else {}
                done, term = rows.get("done"), rows.get("terminated")
                info[done_key] = truncated.clone() if done is None else done.reshape(truncated.shape) | truncated
                info[terminated_key] = torch.zeros_like(truncated) if term is None else term.reshape(truncated.shape)
        if is_init_all is not None:   # every slice start is an init for recurrent modules (:2217-2268)
            is_init = be.gather([rows_of(is_init_all)], flat, n_rows)[0]
            marker = torch.zeros_like(is_init)
            if slice_starts is None:
                slice_starts = torch.arange(num_slices, device=dev) * seq_length
            marker[slice_starts] = True
            info["is_init"] = marker | is_init
        if nd == 2:
            return (index, col), info
        return (index,), info


def _replace_last(key, new: str):
    return new if isinstance(key, str) else (*key[:-1], new)


class PrioritizedSampler(Sampler):
    r"""Prioritized experience replay sampler (Schaul et al. 2015) -- samplers.py:577-1205.

    :math:`P(i) = p_i^\alpha / \sum_j p_j^\alpha`, importance weight
    :math:`w_i = (p_i^\alpha / \min_j p_j^\alpha)^{-\beta}`.

    Args:
        max_capacity (int): maximum capacity of the buffer.
        alpha (float): prioritisation exponent (0 = uniform).
        beta (float): importance-sampling exponent.
        eps (float): added to priorities so that none is zero. Defaults to 1e-8.
        dtype (torch.dtype): tree dtype, ``torch.float`` (default) or ``torch.double``.
        reduction (str): how multi-dim priorities are reduced: "max", "min", "median" or "mean".
        max_priority_within_buffer (bool): track the max priority among the items currently stored
            instead of the max ever seen.
        device: device holding the trees.  ``None``: the storage's CUDA device at first use.

    Extra keyword (not in the reference):
        semantics ("cpu" | "cuda"): which reference tree the sampling arithmetic reproduces when the two
            differ.  "cpu" (default) = ``SumSegmentTree`` on the host: ``query(0, len)`` returns the root when
            ``len >= size`` and zero-priority leaves are walked past (samplers.py:935-943); "cuda" = the
            reference's CUDA port, which always walks the query and has no back-off.
    """

    def __init__(self, max_capacity: int, alpha: float, beta: float, eps: float = 1e-8,
                 dtype: torch.dtype = torch.float, reduction: str = "max", max_priority_within_buffer: bool = False,
                 device=None, *, semantics: str = "cpu") -> None:
        if alpha < 0:
            raise ValueError(f"alpha must be greater or equal than 0, got alpha={alpha}")
        if beta < 0:
            raise ValueError(f"beta must be greater or equal to 0, got beta={beta}")
        if semantics not in ("cpu", "cuda"):
            raise ValueError("semantics must be 'cpu' or 'cuda'")
        self._max_capacity = int(max_capacity)
        self._alpha = alpha
        self._beta = beta
        self._eps = eps
        self.reduction = reduction
        self.dtype = dtype
        self._max_priority_within_buffer = max_priority_within_buffer
        self._device = torch.device(device) if device is not None else None
        self._semantics = semantics
        self._sum_tree = None
        self._min_tree = None
        #: when set to True, ``sample`` records ``index_ready`` (a CUDA event) right after the tree kernel, so work
        #: that only needs the sampled indices -- typically ``update_priority`` on a side stream -- can overlap
        #: with the storage gather that follows on the sampling stream.
        self.record_index_event = False
        self.index_ready = None
        #: when set to True, each ``sample`` also draws the uniforms of the NEXT one (same generator calls, same order)
        #: so that the RNG kernel is not the first thing a sample has to wait for; only worth it for captured /
        #: multi-stream steps, and only if nothing else consumes the generator between two samples.
        self.predraw = False
        self._u_next = None
        self._u_ready = 0
        self._has_max_priority = False
        self._max_priority_index = None
        if self._device is not None:
            self._init()

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(alpha={self._alpha}, beta={self._beta}, eps={self._eps}, "
                f"reduction={self.reduction})")

    # ---- properties --------------------------------------------------------------------------------
    @property
    def max_size(self) -> int:
        return self._max_capacity

    @property
    def device(self) -> torch.device | None:
        if self._sum_tree is not None:
            return self._sum_tree.device
        return self._device

    @property
    def alpha(self):
        return self._alpha

    @alpha.setter
    def alpha(self, value):
        self._alpha = value

    @property
    def beta(self):
        return self._beta

    @beta.setter
    def beta(self, value):
        self._beta = value

    def __getstate__(self):
        import multiprocessing.context as mpc

        if mpc.get_spawning_popen() is not None:
            raise RuntimeError(
                f"Samplers of type {type(self)} cannot be shared between processes. "
                "Use TensorDictPrioritizedReplayBuffer(sync=False) instead: the writer process gets a uniform "
                "sampler and the learner keeps a local prioritized sampler.")
        return super().__getstate__()

    # ---- tree lifetime -----------------------------------------------------------------------------
    def _maybe_init_from_storage(self, storage: Storage | None) -> None:
        # trees follow the storage's device when no explicit device was given (samplers.py:758-775)
        if self._sum_tree is not None:
            return
        device = self._device
        if device is None and storage is not None:
            sd = getattr(storage, "device", None)
            if sd is not None and sd != "auto":
                device = torch.device(sd)
        if device is None:
            raise RuntimeError(
                "PrioritizedSampler needs a device for its HBM-resident trees: pass device=... or use it with a "
                "storage that has one.")
        self._device = device
        self._init()

    def _init(self) -> None:
        if self.dtype in (torch.float, torch.float32):
            sum_cls, min_cls, dt = SumSegmentTreeFp32, MinSegmentTreeFp32, torch.float32
        elif self.dtype in (torch.double, torch.float64):
            sum_cls, min_cls, dt = SumSegmentTreeFp64, MinSegmentTreeFp64, torch.float64
        else:
            raise NotImplementedError(f"dtype {self.dtype} not supported by PrioritizedSampler")
        # both heaps live in ONE allocation so that a single L2 access-policy window can keep them resident
        cap2 = 2 * ops.backend().tree_capacity(self._max_capacity)
        self._tree_buf = torch.empty(2 * cap2, dtype=dt, device=self._device)
        self._sum_tree = sum_cls(self._max_capacity, self._device, out=self._tree_buf[:cap2])
        self._min_tree = min_cls(self._max_capacity, self._device, out=self._tree_buf[cap2:])
        dev = self._sum_tree.device
        # running max of the RAW priorities ever passed to update_priority (samplers.py:1054-1075), kept on
        # the device so that no call has to synchronise; -inf until the first update.
        self._max_priority_buf = torch.full((1,), float("-inf"), dtype=torch.float32, device=dev)
        self._has_max_priority = False
        self._max_priority_index = None
        from .storages import DeferredStatus

        self._status_mirror = DeferredStatus(dev)    # asynchronous host mirror: errors surface without a sync
        self._status = self._status_mirror.word
        self._status_calls = 0
        self._workspace = None
        self._range_ticket = None
        self._epoch = 0

    def _empty(self) -> None:
        if self._device is not None:
            self._init()

    def pin_l2(self, stream=None) -> int:
        """Keep both trees resident in L2 for kernels launched on / captured from ``stream`` (default: the current
        stream).  Returns the MiB set aside (+1), or 0 when the device has no persisting L2."""
        if self._sum_tree is None:
            raise RuntimeError("the trees do not exist yet (no device known)")
        return ops.backend().l2_persist(self._tree_buf, stream)

    def _tree_epoch(self) -> int:
        self._epoch += 1
        if self._epoch >= 0xFFFFFFFF:
            if self._workspace is not None:
                self._workspace.zero_()
            self._epoch = 1
        return self._epoch

    def _tree_workspace(self, n: int):
        if self._workspace is None:  # ticket + per-cluster item lists (<= 8192 items) and stamps (larger batches)
            self._workspace = ops.backend().tree_workspace(self._max_capacity, self._sum_tree.device)
        return self._workspace

    # ---- max-priority bookkeeping ------------------------------------------------------------------
    @property
    def _max_priority(self) -> tuple:
        if not self._has_max_priority:
            return (None, None)
        return (self._max_priority_buf[0], self._max_priority_index)

    @property
    def default_priority(self):
        # (max_priority + eps) ** alpha, max_priority = 1 before any update (samplers.py:886-893).
        # NB: mark_update feeds this through update_priority, which applies (. + eps) ** alpha AGAIN.
        first = (1 + self._eps) ** self._alpha
        if not self._has_max_priority:
            return first
        mx = self._max_priority_buf[0]
        # (the running max is still -inf if every index of the updates so far was a "skip" marker: the reference
        # returns early then and keeps max_priority = None, samplers.py:1040-1052)
        return torch.where(mx > float("-inf"), (mx + self._eps) ** self._alpha, mx.new_full((), first))

    # ---- sample (samplers.py:895-956) --------------------------------------------------------------
    def sample(self, storage: Storage, batch_size: int) -> tuple[Any, dict]:
        self._maybe_init_from_storage(storage)
        length = len(storage)
        if length == 0:
            raise RuntimeError(_EMPTY_STORAGE_ERROR)
        dev = self._sum_tree.device
        self._poll_status()
        u = self._draw(batch_size, dev)
        index, weight = ops.backend().per_sample(
            self._sum_tree.values, self._min_tree.values, self._max_capacity, self._sum_tree.capacity, length, u,
            self._beta, self._semantics == "cpu", status=self._status)
        if self.record_index_event and dev.type == "cuda":
            if self.index_ready is None:
                self.index_ready = torch.cuda.Event()
            self.index_ready.record(torch.cuda.current_stream(dev))
        if self.predraw:
            # the NEXT call's uniforms, drawn now: the same torch.rand calls in the same order, but off the critical
            # path of the next sample (it only has to wait for the priority write-back, not for an RNG kernel first).
            # Issued AFTER the index event: a write-back waiting for the indices must not wait for this kernel too.
            torch.rand(batch_size, device=dev, generator=self._rng, dtype=u.dtype, out=self._u_next)
            self._u_ready = batch_size
        if storage.ndim > 1:
            index = unravel_index(index, storage.shape)
        return index, {"priority_weight": weight}

    def _draw(self, batch_size: int, dev) -> torch.Tensor:
        """``torch.rand(batch_size, generator)`` -- the reference's call (samplers.py:918) -- or, with ``predraw``, the
        values that call produced at the end of the previous ``sample``.  The buffer is persistent so that a captured
        step reads and refills the same memory on every replay."""
        dtype = self._sum_tree._dtype
        if not self.predraw:
            return torch.rand(batch_size, device=dev, generator=self._rng, dtype=dtype)
        if self._u_next is None or self._u_next.numel() != batch_size or self._u_next.device != dev:
            self._u_next = torch.empty(batch_size, device=dev, dtype=dtype)
            self._u_ready = 0
        if self._u_ready != batch_size:      # first call, or another batch size: draw now
            torch.rand(batch_size, device=dev, generator=self._rng, dtype=dtype, out=self._u_next)
        # ONE buffer: the sample kernel reads it, the refill that follows on the same stream overwrites it (stream
        # order keeps the two apart), and a captured step therefore reads what its previous replay drew
        return self._u_next

    #: how often ``sample`` starts an asynchronous copy of the device status word (every call would add a small D2H)
    status_check_every = 16

    def _poll_status(self) -> None:
        """The deferred-error contract: the CPU reference raises "non-positive p_sum / p_min" and "Failed to find a
        suitable index" inside the offending ``sample`` (samplers.py:910-914, :940-941).  Here the kernel ORs a bit into a
        device word; every ``status_check_every``-th eager ``sample`` mirrors it to pinned host memory asynchronously and
        a LATER ``sample`` raises once the copy has landed -- no synchronisation on the hot path (``check_status()``
        synchronises and raises at once; captured steps never poll)."""
        mirror = getattr(self, "_status_mirror", None)
        if mirror is None or (mirror._cuda and torch.cuda.is_current_stream_capturing()):
            return    # (an event query would invalidate a capture)
        bits = mirror.poll()
        if bits:
            self._raise_status(bits)
        self._status_calls += 1
        if self._status_calls % self.status_check_every == 0:
            mirror.arm()

    @staticmethod
    def _raise_status(st: int) -> None:
        if st & ops.STATUS_NONPOS_PSUM:
            raise RuntimeError("non-positive p_sum")
        if st & ops.STATUS_NONPOS_PMIN:
            raise RuntimeError("non-positive p_min")
        if st & ops.STATUS_BACKOFF_FAIL:
            raise RuntimeError("Failed to find a suitable index")

    def check_status(self) -> None:
        """Synchronise and raise what the CPU reference raises eagerly (samplers.py:910-914,940-941)."""
        st = int(self._status.item())
        self._status.zero_()
        if st & ops.STATUS_NONPOS_PSUM:
            raise RuntimeError("non-positive p_sum")
        if st & ops.STATUS_NONPOS_PMIN:
            raise RuntimeError("non-positive p_min")
        if st & ops.STATUS_BACKOFF_FAIL:
            raise RuntimeError("Failed to find a suitable index")

    def add(self, index) -> None:
        super().add(index)
        self._maybe_erase_max_priority(index)

    def extend(self, index) -> None:
        super().extend(index)
        self._maybe_erase_max_priority(index)

    def _maybe_erase_max_priority(self, index) -> None:
        # only meaningful with max_priority_within_buffer (samplers.py:840-884): forget the max when the item
        # that held it is overwritten.  Device-side comparisons would need a sync; like the reference's CUDA
        # branch (:876-878) we simply drop the max.
        if not self._max_priority_within_buffer or not self._has_max_priority:
            return
        self._has_max_priority = False
        self._max_priority_buf.fill_(float("-inf"))
        self._max_priority_index = None

    # ---- update_priority (samplers.py:966-1091) ----------------------------------------------------
    @torch.no_grad()
    def update_priority(self, index, priority, *, storage: Storage | None = None, index_base: int = 0,
                        index_limit: int = -1) -> None:
        """Updates the priority of the data pointed by the index.

        Args:
            index (int or torch.Tensor): indexes of the priorities to be updated.
            priority (Number or torch.Tensor): new priorities of the indexed elements.

        Keyword Args:
            storage (Storage, optional): needed to map N-d indices to the trees' flat index.
            index_base, index_limit (int): (not in the reference) ``index_base`` is subtracted from every index
                and results outside ``[0, index_limit)`` are skipped inside the kernel -- how a shard of the
                capacity-sharded buffer consumes GLOBAL indices without a host round trip.
        """
        self._maybe_init_from_storage(storage)
        dev = self._sum_tree.device
        priority = torch.as_tensor(priority, device=dev).detach()
        index = torch.as_tensor(index, dtype=torch.long, device=dev)
        if priority.numel() > 1 and priority.shape != index.shape:
            try:
                priority = priority.reshape(index.shape[:1])
            except Exception as err:
                raise RuntimeError(
                    "priority should be a number or an iterable of the same "
                    f"length as index. Got priority of shape {priority.shape} and index {index.shape}.") from err
        elif priority.numel() <= 1:
            priority = priority.squeeze()
        if index.ndim == 0:
            index = index.view(1)
            if priority.ndim == 0:
                priority = priority.view(1)
        if index.ndim > 1:
            if storage is None:
                raise RuntimeError(
                    "storage should be provided to Sampler.update_priority when the storage has more "
                    "than one dimension.")
            try:
                shape = storage.shape
            except AttributeError:
                raise AttributeError(
                    "Could not retrieve the storage shape. If your storage is not a TensorStorage subclass "
                    "or its shape isn't accessible via the shape attribute, submit an issue on GitHub.")
            mult = torch.ones(index.shape[-1], dtype=torch.long, device=dev)
            for d in range(index.shape[-1] - 2, -1, -1):
                mult[d] = mult[d + 1] * shape[d + 1]
            index = (index * mult).sum(-1)
        if index.numel() == 0:
            return
        index = index.reshape(-1)
        priority = priority.reshape(-1)
        # negative indices (MaxValueWriter's "do not write" marker, :1040-1052) are skipped by the kernel
        tree_dtype = self._sum_tree._dtype
        if tree_dtype == torch.float32:
            priority = priority.to(torch.float32)
            n = index.numel()
            # batches up to 8192 items are ONE launch (one cluster up to 1024 items, up to 16 above); larger ones use an
            # epoch-stamped scatter whose epoch would be frozen into a captured graph, so under capture they are applied
            # as consecutive chunks of <= 8192 instead (input order is preserved: "the last duplicate wins" still holds)
            step = 8192 if (n > 8192 and dev.type == "cuda" and torch.cuda.is_current_stream_capturing()) else n
            for lo in range(0, n, step):
                pr = priority if priority.numel() == 1 else priority[lo:lo + step]
                ops.backend().per_update(self._sum_tree.values, self._min_tree.values, self._sum_tree.capacity,
                                         index[lo:lo + step], pr, self._alpha, self._eps, self._max_priority_buf,
                                         self._tree_workspace(min(step, n - lo)), self._tree_epoch(), index_base,
                                         index_limit)
        else:
            if index_base or index_limit >= 0:
                index = index - index_base
                lim = self._sum_tree.capacity if index_limit < 0 else index_limit
                index = torch.where((index >= 0) & (index < lim), index, index.new_full((), -1))
            valid = index >= 0
            pmax = torch.where(valid, priority.expand_as(index), priority.new_full((), float("-inf"))).max()
            self._max_priority_buf.copy_(torch.maximum(self._max_priority_buf[0], pmax.to(torch.float32)).view(1))
            leaf = torch.pow(priority.to(tree_dtype) + self._eps, self._alpha)
            ops.backend().tree_update(self._sum_tree.values, self._min_tree.values, self._sum_tree.capacity, index,
                                      leaf, self._tree_workspace(index.numel()), self._tree_epoch())
        self._has_max_priority = True
        if self._max_priority_within_buffer:
            # O(N) rescan of the leaves, as the reference does (samplers.py:1079-1091) -- one reduction here
            leaves = self._sum_tree.values[self._sum_tree.capacity:self._sum_tree.capacity + self._max_capacity]
            maxval, maxidx = leaves.max(0)
            self._max_priority_buf.copy_(maxval.to(torch.float32).view(1))
            self._max_priority_index = maxidx

    def mark_update(self, index, *, storage: Storage | None = None) -> None:
        self._maybe_init_from_storage(storage)
        self.update_priority(index, self.default_priority, storage=storage)

    def _range_update(self, modulo: int, *, storage: Storage | None = None):
        """``mark_update`` of a writer batch -- slots (cursor + arange(n)) % modulo -- as kernel arguments
        (``ops.RangeUpdate``): default priority, its second pow, the running max and the tree write all happen in the
        range kernel (csrc/tree_range.cuh), alone (``mark_update_range``) or fused with the row write
        (``rlb_extend``).  ``None`` when this sampler needs the general path (fp64 trees, whose default priority is
        computed in double, or the O(N) max rescan of ``max_priority_within_buffer``)."""
        self._maybe_init_from_storage(storage)
        if self._max_priority_within_buffer or self._sum_tree._dtype != torch.float32 or modulo > self._max_capacity:
            return None
        if self._range_ticket is None:
            self._range_ticket = torch.zeros(1, dtype=torch.int32, device=self._sum_tree.device)
        # has_max=True: the kernel itself checks whether the running max is still -inf (no priority seen yet)
        rng = ops.RangeUpdate(self._sum_tree.values, self._min_tree.values, self._sum_tree.capacity,
                              ops.RANGE_DEFAULT, alpha=self._alpha, eps=self._eps,
                              first_default=(1 + self._eps) ** self._alpha, has_max=True,
                              max_buf=self._max_priority_buf, ticket=self._range_ticket)
        self._has_max_priority = True   # the kernel about to be launched publishes the new running max
        return rng

    def mark_update_range(self, start: int, n: int, modulo: int, *, storage: Storage | None = None) -> None:
        """``mark_update(arange(start, start + n) % modulo)`` in one small launch (not in the reference API)."""
        rng = self._range_update(modulo, storage=storage)
        if rng is None:
            dev = self._sum_tree.device
            return self.mark_update(torch.arange(start, start + n, device=dev) % modulo, storage=storage)
        ops.backend().tree_update_range(rng, start, n, modulo)

    # ---- (de)serialisation -------------------------------------------------------------------------
    def state_dict(self) -> dict:
        mp = self._max_priority
        return {
            "_alpha": self._alpha,
            "_beta": self._beta,
            "_eps": self._eps,
            "_max_priority": (None if mp[0] is None else float(mp[0]),
                              None if mp[1] is None else int(mp[1])),
            "_sum_tree": deepcopy(self._sum_tree),
            "_min_tree": deepcopy(self._min_tree),
        }

    def load_state_dict(self, state_dict: dict) -> None:
        self._alpha = state_dict["_alpha"]
        self._beta = state_dict["_beta"]
        self._eps = state_dict["_eps"]
        st, mt = state_dict["_sum_tree"], state_dict["_min_tree"]
        if self._sum_tree is None:
            self._device = st.device
            self._init()
        self._sum_tree.load_leaves(st.dump_leaves())
        self._min_tree.load_leaves(mt.dump_leaves())
        self._set_max_priority(state_dict["_max_priority"])

    def _set_max_priority(self, mp) -> None:
        val, idx = mp if mp is not None else (None, None)
        self._has_max_priority = val is not None
        self._max_priority_buf.fill_(float("-inf") if val is None else float(val))
        self._max_priority_index = idx

    def dumps(self, path) -> None:
        """Same on-disk layout as the reference (samplers.py:1120-1163): the LEAVES of both trees as float64
        ``sumtree.memmap`` / ``mintree.memmap`` + ``sampler_metadata.json`` -- written with one D2H copy per
        tree instead of one pybind call per element."""
        path = Path(path).absolute()
        path.mkdir(exist_ok=True, parents=True)
        if self._sum_tree is None:
            raise RuntimeError("cannot dump a PrioritizedSampler whose trees were never created")
        for name, tree in (("sumtree.memmap", self._sum_tree), ("mintree.memmap", self._min_tree)):
            arr = np.memmap(path / name, dtype=np.float64, mode="w+", shape=(self._max_capacity,))
            arr[:] = tree.dump_leaves().to(torch.float64).cpu().numpy()
            arr.flush()
        mp = self._max_priority
        with open(path / "sampler_metadata.json", "w") as file:
            json.dump({"_alpha": float(self._alpha), "_beta": float(self._beta), "_eps": float(self._eps),
                       "_max_priority": [None if mp[0] is None else float(mp[0]),
                                         None if mp[1] is None else float(mp[1])],
                       "_max_capacity": float(self._max_capacity)}, file)

    def loads(self, path) -> None:
        path = Path(path).absolute()
        with open(path / "sampler_metadata.json") as file:
            metadata = json.load(file)
        self._alpha = metadata["_alpha"]
        self._beta = metadata["_beta"]
        self._eps = metadata["_eps"]
        cap = int(metadata["_max_capacity"])
        if cap != self._max_capacity:
            raise RuntimeError(
                f"max capacity of loaded metadata ({cap}) differs from self._max_capacity ({self._max_capacity}).")
        if self._sum_tree is None:
            if self._device is None:
                raise RuntimeError("pass device=... to PrioritizedSampler before loads()")
            self._init()
        for name, tree in (("sumtree.memmap", self._sum_tree), ("mintree.memmap", self._min_tree)):
            arr = np.memmap(path / name, dtype=np.float64, mode="r", shape=(self._max_capacity,))
            tree.load_leaves(torch.from_numpy(np.array(arr)))
        mp = metadata["_max_priority"]
        self._set_max_priority((mp[0], None if mp[1] is None else int(mp[1])))


class SliceSamplerWithoutReplacement(SliceSampler, SamplerWithoutReplacement):
    """Samples slices of data along the first dimension, without replacement over TRAJECTORIES (samplers.py:2303-2573).

    Every sweep visits each stored trajectory once (``torch.randperm(n_trajectories, generator)``, or storage order with
    ``shuffle=False``); the position of the slice inside its trajectory is still drawn uniformly.  Keyword arguments are
    those of :class:`SliceSampler` (``cache_values`` is always on, as in the reference) plus ``drop_last`` / ``shuffle`` of
    :class:`SamplerWithoutReplacement`.  With ``strict_length`` the trajectories of a batch that are too short are
    dropped from it, so a batch may hold fewer slices than asked (:2010-2025).
    """

    def __init__(self, *, num_slices: int | None = None, slice_len: int | None = None, drop_last: bool = False,
                 end_key=None, traj_key=None, ends: torch.Tensor | None = None, trajectories: torch.Tensor | None = None,
                 truncated_key=("next", "truncated"), strict_length: bool = True, shuffle: bool = True, compile=False,
                 use_gpu=False):
        SliceSampler.__init__(self, num_slices=num_slices, slice_len=slice_len, end_key=end_key, traj_key=traj_key,
                              cache_values=True, truncated_key=truncated_key, strict_length=strict_length, ends=ends,
                              trajectories=trajectories, compile=compile, use_gpu=use_gpu)
        SamplerWithoutReplacement.__init__(self, drop_last=drop_last, shuffle=shuffle)
        self._storage_len_buffer = 0

    def __repr__(self) -> str:
        perc = len(self._sample_list) / self.len_storage * 100 if self._sample_list is not None and self.len_storage else 0
        return (f"{self.__class__.__name__}(num_slices={self.num_slices}, slice_len={self.slice_len}, "
                f"end_key={self.end_key}, traj_key={self.traj_key}, truncated_key={self.truncated_key}, "
                f"strict_length={self.strict_length},{perc}% sampled)")

    def _empty(self) -> None:
        self._cache = {}
        SamplerWithoutReplacement._empty(self)

    def _storage_len(self, storage) -> int:
        return self._storage_len_buffer      # the sweep is over trajectories, not steps

    def state_dict(self) -> dict:
        return SamplerWithoutReplacement.state_dict(self)

    def load_state_dict(self, state_dict: dict) -> None:
        SamplerWithoutReplacement.load_state_dict(self, state_dict)

    def dumps(self, path) -> None:
        SamplerWithoutReplacement.dumps(self, path)

    def loads(self, path) -> None:
        SamplerWithoutReplacement.loads(self, path)

    def _plan(self, storage, seq_length: int, num_slices: int):
        keep, self.strict_length = self.strict_length, False          # the sweep is over ALL trajectories of the ring
        try:
            table, n_all, n_long = self._table(storage, seq_length)
        finally:
            self.strict_length = keep
        self._storage_len_buffer = n_all
        traj, _ = SamplerWithoutReplacement.sample(self, storage, num_slices)           # :2540-2541
        if not self.strict_length:
            return table, n_all, n_long < n_all, traj
        if n_long == n_all:
            return table, n_all, False, traj
        if n_long == 0:
            raise RuntimeError("Did not find a single trajectory with sufficient length "
                               f"(required={seq_length}, trajectories={n_all}).")
        # drop the short trajectories of this batch and renumber the others within the long-enough ones (:2010-2025)
        valid = table[2, :n_all] >= seq_length
        picked = torch.zeros(n_all, dtype=torch.bool, device=valid.device)
        picked[traj] = True
        traj = picked[valid].nonzero().squeeze(-1)
        if not traj.numel():
            raise RuntimeError("None of the provided indices pointed to a trajectory of sufficient length. Consider "
                               "using strict_length=False for the sampler instead.")
        return table[:, :n_all][:, valid].contiguous(), n_long, False, traj


class PrioritizedSliceSampler(SliceSampler, PrioritizedSampler):
    """Samples slices of data along the first dimension with prioritized START steps (samplers.py:2575-3028).

    The start of every slice is drawn with probability proportional to its priority among the steps from which a whole
    slice fits inside the trajectory; the slice then runs ``slice_len`` steps forward and every step carries the start's
    importance weight.  Constructor arguments are those of :class:`PrioritizedSampler` followed by the keyword arguments
    of :class:`SliceSampler` (``span=False`` only; 1-d storages).  With ``strict_length=False`` the first step of every
    trajectory stays a legal start however short the trajectory is, and a slice stops where its trajectory stops, so
    slices may be shorter than ``slice_len`` and the batch smaller than asked (:2863-2871, :2919-2951).

    The reference forbids bad starts by zeroing their leaves in the sum tree before each draw and writing them back
    afterwards -- two tree updates of ``n_trajectories * (slice_len - 1)`` items per sample, done index by index on the
    host (:2910-2918).  Here the draw comes from a masked copy: one device copy of the sum tree's leaves,
    ``rlb_slice_mask_starts`` (zero the last ``slice_len - 1`` leaves of every trajectory of the ``rlb_traj_table``),
    ``rlb_tree_rebuild`` and the ordinary ``rlb_per_sample`` with the untouched min tree.  A tree built from the masked
    leaves has, node for node, the values the reference's zero-and-recompute produces, so the sampled starts are the
    reference's for the same uniform draws.
    """

    def __init__(self, max_capacity: int, alpha: float, beta: float, eps: float = 1e-8,
                 dtype: torch.dtype = torch.float, reduction: str = "max", *, num_slices: int | None = None,
                 slice_len: int | None = None, end_key=None, traj_key=None, ends: torch.Tensor | None = None,
                 trajectories: torch.Tensor | None = None, cache_values: bool = False,
                 truncated_key=("next", "truncated"), strict_length: bool = True, compile=False, span=False,
                 max_priority_within_buffer: bool = False, device=None, semantics: str = "cpu"):
        SliceSampler.__init__(self, num_slices=num_slices, slice_len=slice_len, end_key=end_key, traj_key=traj_key,
                              cache_values=cache_values, truncated_key=truncated_key, strict_length=strict_length,
                              ends=ends, trajectories=trajectories, compile=compile, span=span)
        PrioritizedSampler.__init__(self, max_capacity=max_capacity, alpha=alpha, beta=beta, eps=eps, dtype=dtype,
                                    reduction=reduction, max_priority_within_buffer=max_priority_within_buffer,
                                    device=device, semantics=semantics)
        self._masked = None

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(num_slices={self.num_slices}, slice_len={self.slice_len}, "
                f"end_key={self.end_key}, traj_key={self.traj_key}, truncated_key={self.truncated_key}, "
                f"strict_length={self.strict_length}, alpha={self._alpha}, beta={self._beta}, eps={self._eps})")

    # the priority bookkeeping is PrioritizedSampler's, the table cache SliceSampler's
    def mark_update(self, index, *, storage: Storage | None = None) -> None:
        return PrioritizedSampler.mark_update(self, index, storage=storage)

    # add / extend: SliceSampler's (drop the cached table), which chain to PrioritizedSampler's through super()

    def _empty(self) -> None:
        SliceSampler._empty(self)
        PrioritizedSampler._empty(self)

    def dumps(self, path) -> None:
        PrioritizedSampler.dumps(self, path)

    def loads(self, path) -> None:
        PrioritizedSampler.loads(self, path)

    def state_dict(self) -> dict:
        return PrioritizedSampler.state_dict(self)

    def load_state_dict(self, state_dict: dict) -> None:
        PrioritizedSampler.load_state_dict(self, state_dict)

    def sample(self, storage: Storage, batch_size: int) -> tuple[Any, dict]:
        if storage.ndim != 1:
            raise NotImplementedError("PrioritizedSliceSampler on the B200 engine supports 1-d storages only")
        self._maybe_init_from_storage(storage)
        length = len(storage)
        if length == 0:
            raise RuntimeError(_EMPTY_STORAGE_ERROR)
        seq_length, num_slices = self._adjusted_batch_size(batch_size)
        # every trajectory of the ring (no length filter: short ones are masked out whole)
        keep, self.strict_length = self.strict_length, False
        try:
            table, n_all, _ = self._table(storage, seq_length)
        finally:
            self.strict_length = keep
        be = ops.backend()
        st, mt = self._sum_tree, self._min_tree
        if self._masked is None or self._masked.shape != st.values.shape:
            self._masked = torch.empty_like(st.values)
        cap = st.capacity
        self._masked[cap:].copy_(st.values[cap:])                                        # :2910 (vals = tree[idx])
        # strict: the last slice_len - 1 steps of a trajectory cannot start a slice (all of a shorter one); loose: the
        # same but never its FIRST step (:2863-2871 drops the starts from the candidates) -- i.e. one step fewer
        tail = table[2] if self.strict_length else table[2] - 1
        be.slice_mask_starts(self._masked, cap, table[1], tail, n_all, seq_length, storage.shape[0])   # :2911
        be.tree_rebuild(self._masked, cap, False)
        dev = st.device
        u = torch.rand(num_slices, device=dev, generator=self._rng, dtype=st._dtype)       # PrioritizedSampler.sample
        starts, weight = be.per_sample(self._masked, mt.values, self._max_capacity, cap, length, u, self._beta,
                                       self._semantics == "cpu", status=self._status)
        if self.record_index_event and dev.type == "cuda":
            if self.index_ready is None:
                self.index_ready = torch.cuda.Event()
            self.index_ready.record(torch.cuda.current_stream(dev))
        seq = None
        if not self.strict_length:
            # :2919-2951 -- a slice ends with its trajectory: the stop that follows the start (the table is ordered by
            # stop; past the last one the trajectory wraps and ends at the first stop of the ring)
            stops = table[1, :n_all]
            j = torch.searchsorted(stops, starts)
            ring = storage.shape[0]
            stop_after = torch.where(j < n_all, stops[j.clamp_max(n_all - 1)], stops[0] + ring)
            seq = (stop_after - starts + 1).clamp_max(seq_length)
        info: dict = {"priority_weight": weight.repeat_interleave(seq_length if seq is None else seq)}   # :2969-2971
        # expansion of the starts (:2963-2966), truncated markers and the stored flags of the sampled steps in one
        # rlb_slice_index launch: every start is a one-entry "trajectory" of exactly its slice's steps, offset 0
        contents = _contents(storage)
        done_all = term_all = None
        if self.truncated_key is not None:
            done_key = _replace_last(self.truncated_key, "done")
            terminated_key = _replace_last(self.truncated_key, "terminated")
            done_all, term_all = contents.get(done_key, None), contents.get(terminated_key, None)
        one_byte = lambda t: t is None or (t.element_size() == 1 and t.numel() == t.shape[0] and t.is_contiguous())
        fused = self.truncated_key is not None and one_byte(done_all) and one_byte(term_all)
        kw = dict(flags=(done_all, term_all)) if fused else {}
        arange, zeros = torch.arange(num_slices, device=dev), torch.zeros(num_slices, device=dev)
        if seq is None:
            out = be.slice_index(starts, torch.full_like(starts, seq_length), num_slices, arange, zeros, seq_length,
                                 storage.shape[0], **kw)
        else:
            ends_at = seq.cumsum(0)
            out = be.slice_index(starts, seq, num_slices, arange, zeros, seq_length, storage.shape[0], variable=True,
                                 out_offset=ends_at - seq, total=int(ends_at[-1]), **kw)      # data-dependent batch size
        index, truncated = out[0], out[1]
        if self.truncated_key is not None:
            info[self.truncated_key] = truncated
            if fused:
                info[done_key], info[terminated_key] = out[4], out[5]
            else:
                have = {k: v for k, v in (("done", done_all), ("terminated", term_all)) if v is not None}
                rows = dict(zip(have, be.gather(list(have.values()), index, length))) if have else {}
                done, term = rows.get("done"), rows.get("terminated")
                info[done_key] = truncated.clone() if done is None else done.reshape(truncated.shape) | truncated
                info[terminated_key] = torch.zeros_like(truncated) if term is None else term
        return (index,), info
