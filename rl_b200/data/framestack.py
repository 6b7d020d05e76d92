"""De-duplicated frame-stack storage.

The reference has no counterpart: a ``TensorStorage`` keeps whatever it is handed (storages.py:1028-1096), so an Atari DQN
buffer holds, per transition, the 4-frame observation stack AND the 4-frame next-observation stack -- 8 frames = 56 448 B of
which 7 repeat frames of the neighbouring transitions.  ``FrameStackStorage`` is a drop-in storage for the same
transitions (same ``set`` / ``get`` / ``len`` contract, same batch out, bit for bit) that logs every frame ONCE:

  * write (``rlb_framestack_push``, csrc/framestack.cu): per environment stream, an episode's first transition logs its k
    observation frames and its newest next-observation frame, every other transition logs the newest frame only; the
    transition keeps one int64 frame word (environment, log position).
  * read: the stacks are rebuilt by the gather kernel itself (``rlb_gather_ex``, csrc/gather.cu resolve_row) -- frame
    j of ``obs`` is log position p - k + j, frame j of ``next`` is p - k + 1 + j -- in the SAME launch that gathers the other
    leaves.  The k + 1 distinct frames of a transition are read from HBM once (the second use hits L2).

Pixel bytes per transition: ``(1 + (k + 1) / episode_length)`` frames instead of ``2 k`` -- 7.1 KB instead of 56.4 KB for
Atari at the default pool size, i.e. ~7x the transitions per GB of HBM.

The de-duplication relies on the data really being a frame stack of a stream: ``obs[t] == next[t - 1]`` inside an episode
and ``next[t][:-1] == obs[t][1:]``.  ``validate=True`` reads every written batch back and compares (one host sync per
write; for tests and first runs).  Episode starts come from ``init_key`` (``"is_init"``, the reference's ``InitTracker``)
when the data has it, else from ``done_key`` of the previous transition of the same environment.
"""
from __future__ import annotations

import json
from pathlib import Path

import torch

from .. import ops
from .storages import DeferredStatus, LazyTensorStorage, Storage
from .tensordict_lite import TensorDict, _norm_key, is_tensor_collection
from .utils import _is_int

FRAME_WORD_KEY = "_frame_word"


class FrameStackStorage(Storage):
    """HBM-resident storage of frame-stacked transitions that keeps every frame once.

    Args:
        max_size: capacity in transitions.
    Keyword Args:
        n_envs: number of environment streams interleaved in every written batch (1: a single stream).
        batch_layout: ``"env_major"`` -- row ``i`` of a written batch is step ``i % T`` of environment ``i // T`` (a
            ``[E, T]`` collector batch flattened, the reference collectors' layout) -- or ``"time_major"`` (``[T, E]``
            flattened).  Irrelevant for ``n_envs=1``.  Every batch holds the same number of consecutive steps of every
            environment, in stream order.
        obs_key, next_key: the two stacked leaves, ``[n, k, *frame]`` each.
        init_key, done_key: see the module docstring.
        frame_capacity: frames per environment ring.  Default: ``steps + k + k * (steps // min_episode_length + 2)`` with
            ``steps = ceil(max_size / n_envs)`` -- enough for every stored transition as long as episodes are at least
            ``min_episode_length`` steps long on average.  A transition whose frames were overwritten is reported when it
            is sampled (``RuntimeError``), never returned silently.
        materialize: ``True``: ``obs`` and ``next`` come back as two contiguous ``[B, k, *frame]`` tensors like the
            reference's; ``False``: as two overlapping views of ONE ``[B, k + 1, *frame]`` window (5/8 of the bytes).
        validate: read back and compare every write.
    """

    def __init__(self, max_size: int, *, n_envs: int = 1, batch_layout: str = "env_major", obs_key="pixels",
                 next_key=("next", "pixels"), init_key="is_init", done_key=("next", "done"),
                 frame_capacity: int | None = None, min_episode_length: int = 16, device="cuda",
                 materialize: bool = True, validate: bool = False):
        super().__init__(max_size)
        if batch_layout not in ("env_major", "time_major"):
            raise ValueError("batch_layout must be 'env_major' or 'time_major'")
        if n_envs < 1:
            raise ValueError("n_envs must be positive")
        self.n_envs = int(n_envs)
        self.batch_layout = batch_layout
        self.obs_key, self.next_key = _norm_key(obs_key), _norm_key(next_key)
        self.init_key, self.done_key = _norm_key(init_key), _norm_key(done_key)
        self.frame_capacity = frame_capacity
        self.min_episode_length = int(min_episode_length)
        self.materialize = bool(materialize)
        self.validate = bool(validate)
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self._inner = LazyTensorStorage(max_size, device=self.device)
        self._inner.enable_index_check(False)   # this class keeps the status word of its own launches
        self.num_frames = None      # k
        self._pool = None           # [n_envs * ring, *frame]
        self._head = None           # int64 [n_envs]: frames logged per env
        self._last_done = None      # uint8 [n_envs]
        self._ring = None
        self._plans = {}            # materialize? -> (plan, the inner leaves it was built for, keys, kept leaf numbers)
        self._status = None

    # ---- layout ------------------------------------------------------------------------------------
    @property
    def initialized(self) -> bool:
        return self._pool is not None

    @property
    def frame_bytes_per_transition(self) -> float:
        """Pixel bytes this storage spends per transition slot (the reference: ``2 * k`` frames)."""
        return self._pool.numel() * self._pool.element_size() / self.max_size

    def _init_pool(self, obs: torch.Tensor) -> None:
        if obs.ndim < 3:
            raise RuntimeError(f"{self.obs_key}: expected [n, k, *frame] stacks, got shape {tuple(obs.shape)}")
        k, frame = int(obs.shape[1]), tuple(obs.shape[2:])
        steps = -(-self.max_size // self.n_envs)
        ring = self.frame_capacity
        if ring is None:
            ring = steps + k + k * (steps // max(self.min_episode_length, 1) + 2)
        self.num_frames, self._ring = k, int(ring)
        self._pool = torch.empty((self.n_envs * self._ring, *frame), dtype=obs.dtype, device=self.device)
        self._head = torch.zeros(self.n_envs, dtype=torch.int64, device=self.device)
        self._last_done = torch.ones(self.n_envs, dtype=torch.uint8, device=self.device)   # the first step starts an episode

    def _split(self, data):
        """data -> (obs, next, is_init | None, done | None, the other leaves + room for the frame word)."""
        if not is_tensor_collection(data):
            raise RuntimeError("FrameStackStorage stores tensordict-structured transitions")
        if data.batch_dims != 1:
            raise RuntimeError("FrameStackStorage expects a flat batch of transitions (reshape(-1) the collector batch)")
        obs, nxt = data.get(self.obs_key, None), data.get(self.next_key, None)
        if obs is None or nxt is None:
            raise KeyError(f"the data must hold {self.obs_key} and {self.next_key}")
        if obs.shape != nxt.shape or obs.dtype != nxt.dtype:
            raise RuntimeError("the observation and next-observation stacks differ in shape or dtype")
        init, done = data.get(self.init_key, None), data.get(self.done_key, None)
        if init is None and done is None:
            raise KeyError(f"the data must hold {self.init_key} or {self.done_key} to find the episode starts")
        rest = TensorDict({}, data.batch_size)
        for key in data.keys(True, True):
            key = _norm_key(key)
            if key not in (self.obs_key, self.next_key):
                rest.set(key, data.get(key))
        return obs, nxt, init, done, rest

    def _flag(self, t, n):
        if t is None:
            return None
        t = t.to(self.device).reshape(n, -1)
        if t.shape[1] != 1:
            t = t.any(dim=1)
        return t.reshape(n).to(torch.bool).contiguous()

    def _push(self, data):
        obs, nxt, init, done, rest = self._split(data)
        n = obs.shape[0]
        if n % self.n_envs:
            raise RuntimeError(f"a batch of {n} transitions does not divide into {self.n_envs} environments")
        obs, nxt = obs.to(self.device), nxt.to(self.device)
        if not self.initialized:
            self._init_pool(obs)
        if tuple(obs.shape[1:]) != (self.num_frames, *self._pool.shape[1:]) or obs.dtype != self._pool.dtype:
            raise RuntimeError(f"cannot write stacks of shape {tuple(obs.shape[1:])} / {obs.dtype} into a pool of "
                               f"{self.num_frames} x {tuple(self._pool.shape[1:])} / {self._pool.dtype} frames")
        if not obs[0].is_contiguous():
            obs = obs.contiguous()
        if not nxt[0].is_contiguous():
            nxt = nxt.contiguous()
        word = ops.backend().framestack_push(obs, nxt, self._flag(init, n), self._flag(done, n), self._last_done,
                                             self._head, self._pool, self.n_envs,
                                             0 if self.batch_layout == "env_major" else 1, self.num_frames, self._ring)
        rest.set(FRAME_WORD_KEY, word)
        return rest, obs, nxt

    # ---- writes ------------------------------------------------------------------------------------
    def set(self, cursor, data, *, set_cursor: bool = True):
        if not set_cursor:
            raise RuntimeError("FrameStackStorage rows are written in stream order by a writer; in-place row edits would "
                               "break the frame sharing")
        rest, obs, nxt = self._push(data)
        self._inner.set(cursor, rest, set_cursor=True)
        if self.validate:
            self._read_back(cursor, obs, nxt)

    def _fits_range(self, n: int, data) -> bool:
        return 0 < n <= self.max_size and is_tensor_collection(data)

    def _extend_range(self, cursor: int, n: int, data, trees=None) -> None:
        """The writer's modular slot range: frames pushed, then the remaining leaves and the default priorities in the
        fused ``rlb_extend`` launch of the inner storage."""
        rest, obs, nxt = self._push(data)
        self._inner._extend_range(cursor, n, rest, trees)
        if self.validate:
            self._read_back(torch.arange(cursor, cursor + n, device=self.device) % self.max_size, obs, nxt)

    def _read_back(self, cursor, obs, nxt) -> None:
        if isinstance(cursor, slice):
            cursor = torch.arange(cursor.start or 0, cursor.stop, device=self.device)
        elif _is_int(cursor):
            cursor = torch.tensor([cursor], device=self.device)
        got = self.get(torch.as_tensor(cursor, device=self.device, dtype=torch.long))
        if not (torch.equal(got.get(self.obs_key), obs) and torch.equal(got.get(self.next_key), nxt)):
            raise RuntimeError("FrameStackStorage(validate=True): the written stacks are not a frame stack of one stream "
                               "per environment (obs[t] != next[t-1] inside an episode, or a wrong n_envs / batch_layout)")

    # ---- reads -------------------------------------------------------------------------------------
    def _gather_plan(self, materialize: bool):
        """The launch plan: the inner leaves (minus the frame words) as ordinary leaves + the frame pool once per output
        frame -- 2 k of them (both stacks written out) or k + 1 (one window the two stacks are views of)."""
        cached = self._plans.get(materialize)
        if cached is None or cached[1] is not self._inner._leaves or cached[4] is not self._pool:
            inner = self._inner
            keys = [_norm_key(k) for k in inner._spec[1]]
            keep = [i for i, k in enumerate(keys) if k != (FRAME_WORD_KEY,)]
            word = inner._leaves[keys.index((FRAME_WORD_KEY,))]
            k = self.num_frames
            offsets = ([j - k for j in range(k)] + [j - k + 1 for j in range(k)]) if materialize \
                else [j - k for j in range(k + 1)]
            leaves = [inner._leaves[i] for i in keep] + [self._pool] * len(offsets)
            frames = [None] * len(keep) + [(word, self._head, self._ring, off) for off in offsets]
            plan = ops.backend().gather_plan(leaves, frames)
            cached = self._plans[materialize] = (plan, inner._leaves, [keys[i] for i in keep], keep, self._pool)
        return cached[:4]

    # -- the packed form the sharded buffer exchanges: the other leaves + ONE [k + 1, *frame] window per transition
    def _packed_templates(self) -> list:
        _, _, _, keep = self._gather_plan(False)
        return [self._inner._leaves[i] for i in keep] + \
            [torch.empty((1, self.num_frames + 1, *self._pool.shape[1:]), dtype=self._pool.dtype, device=self.device)]

    def _gather_packed(self, index: torch.Tensor, out: list, peer_delta=None, multicast_delta: int = 0) -> None:
        """``out``: one [B, ...] (possibly strided) view per template; the last one is the window."""
        plan, _, _, keep = self._gather_plan(False)
        win = out[-1]
        st = self._status_word()
        plan.run(index, len(self), status=st.word, out=list(out[:len(keep)]) + [win[:, j] for j in range(self.num_frames + 1)],
                 peer_delta=peer_delta, multicast_delta=multicast_delta)
        st.arm()

    def _unpack(self, views: list, batch_size):
        _, _, keys, keep = self._gather_plan(False)
        win, k = views[-1], self.num_frames
        return TensorDict._from_leaves(keys + [self.obs_key, self.next_key], list(views[:len(keep)]) + [win[:, :k], win[:, 1:]],
                                       batch_size)

    def _status_word(self):
        if self._status is None:
            self._status = DeferredStatus(self.device)
        st = self._status
        if not (st._cuda and torch.cuda.is_current_stream_capturing()):
            self._raise(st.poll())
        return st

    @staticmethod
    def _raise(bits: int) -> None:
        if bits & ops.STATUS_FRAME_EVICTED:
            raise RuntimeError("FrameStackStorage: a sampled transition's frames had been overwritten in its environment's "
                               "ring -- episodes are shorter than min_episode_length allows for; raise frame_capacity")
        if bits & ops.STATUS_INDEX_OOB:
            raise IndexError("index out of range in an earlier tensor-indexed read of this storage")

    def check_index_status(self) -> None:
        if self._status is not None:
            self._raise(self._status.check())

    def _gather(self, index: torch.Tensor):
        plan, _, keys, keep = self._gather_plan(self.materialize)
        B, k = index.numel(), self.num_frames
        frame = tuple(self._pool.shape[1:])
        inner = self._inner
        out = [torch.empty((B, *inner._leaves[i].shape[1:]), dtype=inner._leaves[i].dtype, device=self.device)
               for i in keep]
        if self.materialize:
            obs = torch.empty((B, k, *frame), dtype=self._pool.dtype, device=self.device)
            nxt = torch.empty((B, k, *frame), dtype=self._pool.dtype, device=self.device)
            out += [obs[:, j] for j in range(k)] + [nxt[:, j] for j in range(k)]
        else:
            win = torch.empty((B, k + 1, *frame), dtype=self._pool.dtype, device=self.device)
            obs, nxt = win[:, :k], win[:, 1:]
            out += [win[:, j] for j in range(k + 1)]
        st = self._status_word()
        plan.run(index, len(self), status=st.word, out=out)
        st.arm()
        return TensorDict._from_leaves(keys + [self.obs_key, self.next_key], out[:len(keep)] + [obs, nxt], (B,))

    def _signal_view(self):
        """Views of every stored leaf except the two stacks (what the slice samplers look their episode signals up in)."""
        return self._inner.get(slice(None))

    @property
    def _last_cursor(self):
        return self._inner._last_cursor

    def _get_trusted(self, index: torch.Tensor):
        return self.get(index)

    def get(self, index):
        if not self.initialized:
            raise RuntimeError("Cannot get elements out of a non-initialized storage.")
        n = len(self)
        squeeze = False
        if isinstance(index, tuple):
            if len(index) != 1:
                raise RuntimeError("FrameStackStorage is one-dimensional: expected a single index")
            index = index[0]
        if _is_int(index):
            index, squeeze = torch.tensor([index + n if index < 0 else index], device=self.device), True
        elif isinstance(index, slice) or index is None or index is Ellipsis:
            index = torch.arange(n, device=self.device)[index if isinstance(index, slice) else slice(None)]
        else:
            index = torch.as_tensor(index)
            if index.dtype == torch.bool:
                index = index.nonzero().squeeze(-1)
            index = index.to(device=self.device, dtype=torch.long)
        shape = index.shape
        out = self._gather(index.reshape(-1).contiguous())
        if squeeze:
            return out[0]
        return out if len(shape) == 1 else out.reshape(*shape)

    def __len__(self) -> int:
        return len(self._inner)

    def _empty(self) -> None:
        self._inner._empty()
        if self._last_done is not None:
            self._last_done.fill_(1)   # whatever comes next starts an episode

    def contains(self, item) -> bool:
        return self._inner.contains(item)

    # ---- persistence -------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        return {"inner": self._inner.state_dict(), "pool": None if self._pool is None else self._pool.cpu(),
                "head": None if self._head is None else self._head.cpu(),
                "last_done": None if self._last_done is None else self._last_done.cpu(),
                "ring": self._ring, "num_frames": self.num_frames}

    def load_state_dict(self, sd: dict) -> None:
        self._inner.load_state_dict(sd["inner"])
        self._plans = {}
        if sd["pool"] is not None:
            self._ring, self.num_frames = sd["ring"], sd["num_frames"]
            self._pool = sd["pool"].to(self.device)
            self._head = sd["head"].to(self.device)
            self._last_done = sd["last_done"].to(self.device)

    def dumps(self, path) -> None:
        """``<path>/transitions``: the inner TensorStorage in the reference's checkpoint layout (every leaf except the two
        stacks, plus the frame words); ``<path>/frames``: the frame pool and the ring state."""
        from .checkpointers import _write_leaf

        if not self.initialized:
            raise RuntimeError("Cannot save a non-initialized storage.")
        path = Path(path)
        self._inner.dumps(path / "transitions")
        fdir = path / "frames"
        fdir.mkdir(parents=True, exist_ok=True)
        _write_leaf(self._pool, self._pool.shape[0], fdir / "pool.memmap")
        meta = {"ring": self._ring, "num_frames": self.num_frames, "n_envs": self.n_envs,
                "frame_shape": list(self._pool.shape[1:]), "dtype": str(self._pool.dtype),
                "head": self._head.cpu().tolist(), "last_done": self._last_done.cpu().tolist()}
        (fdir / "meta.json").write_text(json.dumps(meta))

    def loads(self, path) -> None:
        from .checkpointers import _STRDTYPE2DTYPE, _read_leaf

        path = Path(path)
        meta = json.loads((path / "frames" / "meta.json").read_text())
        if meta["n_envs"] != self.n_envs:
            raise RuntimeError(f"checkpoint has {meta['n_envs']} environment streams, this storage {self.n_envs}")
        self._inner.loads(path / "transitions")
        self._ring, self.num_frames = int(meta["ring"]), int(meta["num_frames"])
        dt = _STRDTYPE2DTYPE[meta["dtype"]]
        shape = (self.n_envs * self._ring, *meta["frame_shape"])
        self._pool = torch.empty(shape, dtype=dt, device=self.device)
        _read_leaf(path / "frames" / "pool.memmap", shape, dt, self._pool, shape[0])
        self._head = torch.tensor(meta["head"], dtype=torch.int64, device=self.device)
        self._last_done = torch.tensor(meta["last_done"], dtype=torch.uint8, device=self.device)
        self._plans = {}

    def __repr__(self) -> str:
        return (f"FrameStackStorage(max_size={self.max_size}, len={len(self)}, n_envs={self.n_envs}, "
                f"num_frames={self.num_frames}, ring={self._ring}, device={self.device})")
