"""Capacity-sharded prioritized replay over the GPUs of one box.

The reference has no sharded / collective replay buffer (its nearest analogue is ``SamplerEnsemble`` with
``sample_from_all=True``, samplers.py:3111-3118: equal quotas per sub-buffer).  Here slots
``[r*N/W, (r+1)*N/W)`` live on rank ``r`` with their own sum/min trees; ``sample(B)`` draws ``B/W`` rows on
every rank and assembles the global minibatch with ONE all-gather over NVLink:

  * the local gather kernel writes every leaf of the local draw straight into its column of a packed
    ``[B/W, row]`` send buffer (``rlb_gather`` with destination strides) -- no staging copy;
  * the same rows carry the global index, the leaf priority ``p_i`` and the shard's ``(S_r, m_r)``
    (sum and min of its priorities), so importance weights can be finalised identically on every rank after
    the gather, normalised over the WHOLE buffer as the reference normalises over its single buffer:
        w_i = ((p_i / S_r) / min_r'(m_r' / S_r')) ** -beta          (W = 1  ->  (p_i / p_min) ** -beta)
  * transport "nvlink" (default on CUDA when torch's symmetric memory is available): the receive buffers of all
    ranks are ONE symmetric allocation (``torch.distributed._symmetric_memory``), so every rank knows the
    address of its rows in every peer's buffer and the gather kernel's shared-memory stages are stored W times
    -- once locally, W-1 times through NVLink peer memory.  Gather and all-gather are the SAME launch
    (``rlb_gather`` with ``peer_delta``); a signal-pad barrier (one tiny kernel) closes the exchange, and the
    whole step is capturable in a CUDA graph.  Receive buffers are double-buffered: a returned batch stays valid
    until the second next ``sample()``.
  * transport "nccl": ``torch.distributed.all_gather_into_tensor`` moves ``B/W * row`` bytes per rank (also the
    CPU/gloo test path).  Either way the leaves of the returned batch are strided views into the receive buffer.

``update_priority`` takes GLOBAL indices (replicated on every rank, or any subset): each rank rewrites the
ones it owns and skips the rest inside the kernel (negative local index) -- no collective, no sync.
"""
from __future__ import annotations


import torch

from .. import ops
from .replay_buffers import TensorDictPrioritizedReplayBuffer
from .storages import LazyTensorStorage, unflatten_data
from .tensordict_lite import is_tensor_collection


class _NoSymmetricMemory(RuntimeError):
    pass


def _align(x: int, a: int) -> int:
    return -(-x // a) * a


class _PackedLayout:
    """Column layout of one packed transition row: every storage leaf + (global index, p_i, S_r, m_r)."""

    def __init__(self, leaves):
        order = sorted(range(len(leaves)), key=lambda k: -leaves[k][0].numel() * leaves[k].element_size())
        off = 0
        self.cols = [None] * len(leaves)
        for k in order:
            t = leaves[k]
            nbytes = t[0].numel() * t.element_size()
            a = 16 if nbytes >= 16 else max(t.element_size(), 1)
            off = _align(off, a)
            self.cols[k] = (off, nbytes, t.dtype, tuple(t.shape[1:]))
            off += nbytes
        off = _align(off, 8)
        self.meta = off            # int64 global index | f32 p_i | f32 S_r | f32 m_r | pad
        off += 8 + 4 + 4 + 4
        self.row = _align(off, 16)

    def leaf_views(self, buf: torch.Tensor) -> list[torch.Tensor]:
        """buf: uint8 [rows, self.row] -> one strided [rows, *shape] view per leaf."""
        rows = buf.shape[0]
        out = []
        for off, nbytes, dtype, shape in self.cols:
            col = buf[:, off:off + nbytes]
            v = col.view(dtype) if dtype != torch.uint8 else col
            out.append(v.view(rows, *shape) if shape else v.view(rows))
        return out

    def meta_views(self, buf: torch.Tensor):
        m = self.meta
        return (buf[:, m:m + 8].view(torch.int64).view(-1), buf[:, m + 8:m + 12].view(torch.float32).view(-1),
                buf[:, m + 12:m + 16].view(torch.float32).view(-1), buf[:, m + 16:m + 20].view(torch.float32).view(-1))


class ShardedPrioritizedReplayBuffer:
    """One shard of a capacity-sharded ``TensorDictPrioritizedReplayBuffer`` per rank.

    Keyword Args:
        alpha, beta, eps, priority_key: as :class:`TensorDictPrioritizedReplayBuffer`.
        capacity (int): GLOBAL capacity; every rank holds ``ceil(capacity / world_size)`` slots.
        batch_size (int): GLOBAL batch size of :meth:`sample` (must be divisible by the world size).
        device: this rank's device.
        generator: this rank's random generator (seed it with ``seed + rank``).
        process_group: defaults to the global group; ``None`` with an uninitialised torch.distributed runs as a
            single shard (world size 1).
    """

    def __init__(self, *, alpha: float, beta: float, capacity: int, eps: float = 1e-8, priority_key: str = "td_error",
                 batch_size: int | None = None, device="cuda", generator=None, process_group=None,
                 transport: str = "auto"):
        import torch.distributed as dist

        self._dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = process_group
        self.world = self._dist.get_world_size(process_group) if self._dist else 1
        self.rank = self._dist.get_rank(process_group) if self._dist else 0
        self.capacity = int(capacity)
        self.shard_capacity = -(-self.capacity // self.world)
        if batch_size is not None and batch_size % self.world:
            raise ValueError(f"batch_size={batch_size} must be divisible by the world size {self.world}")
        self._batch_size = batch_size
        self.device = torch.device(device)
        self.local = TensorDictPrioritizedReplayBuffer(
            alpha=alpha, beta=beta, eps=eps, priority_key=priority_key,
            storage=LazyTensorStorage(self.shard_capacity, device=self.device),
            batch_size=None if batch_size is None else batch_size // self.world, generator=generator)
        if transport not in ("auto", "nvlink", "nccl"):
            raise ValueError("transport must be 'auto', 'nvlink' or 'nccl'")
        self.transport = transport
        self._last_gidx = None
        self._symm = None        # (symmetric double buffer, handle, peer byte offsets)
        self._parity = 0
        self._layout = None
        self._static = None
        self._send = self._recv = None
        self._bs = None
        self.local_index = None  # local indices of this rank's last draw

    # ---- writes: every rank feeds its own shard (data-parallel collectors) --------------------------------
    def extend(self, data) -> torch.Tensor:
        """Writes ``data`` into this rank's shard; returns the GLOBAL indices of the written slots."""
        local = self.local.extend(data)
        return local + self.rank * self.shard_capacity

    def __len__(self) -> int:
        return len(self.local)

    @property
    def sampler(self):
        return self.local.sampler

    @property
    def storage(self):
        return self.local.storage

    # ---- sample --------------------------------------------------------------------------------------
    # sample() = local_draw() -> exchange() -> finalize().  The three stages are public so that a training step can
    # capture the two compute stages in CUDA graphs and issue the collective eagerly in between.
    def _resolve_batch(self, batch_size):
        if batch_size is None:
            batch_size = self._batch_size
        if batch_size is None:
            raise RuntimeError("batch_size not specified.")
        if batch_size % self.world:
            raise ValueError(f"batch_size={batch_size} must be divisible by the world size {self.world}")
        return batch_size

    def local_draw(self, batch_size: int | None = None, *, static_buffers: bool = False) -> torch.Tensor:
        """Draw ``batch_size / world`` rows from this shard straight into the packed send buffer (returned).

        ``static_buffers=True`` reuses one send / receive buffer pair across calls (needed under CUDA-graph capture;
        the returned batch is then overwritten by the next sample).
        """
        batch_size = self._resolve_batch(batch_size)
        b_loc = batch_size // self.world
        st, smp = self.local.storage, self.local.sampler
        smp._maybe_init_from_storage(st)
        length = len(st)
        if length == 0:
            raise RuntimeError("Cannot sample from an empty storage.")
        be = ops.backend()
        dev = smp._sum_tree.device
        if self._layout is None:
            self._layout = _PackedLayout(st._leaves)
        lay = self._layout
        peers = None
        nv = None
        if self._use_nvlink(dev):
            try:
                nv = self._symmetric_buffers(batch_size, dev)
            except _NoSymmetricMemory:
                nv = None
        if nv is not None:
            buf, hdl, peers = nv
            recv = buf[self._parity]
            send = recv[self.rank * b_loc:(self.rank + 1) * b_loc]  # my rows inside the gathered batch
        elif static_buffers:
            if self._static is None or self._static[0].shape[0] != b_loc:
                self._static = (torch.empty((b_loc, lay.row), dtype=torch.uint8, device=dev),
                                torch.empty((batch_size, lay.row), dtype=torch.uint8, device=dev))
            send, recv = self._static
        else:
            send = torch.empty((b_loc, lay.row), dtype=torch.uint8, device=dev)
            recv = torch.empty((batch_size, lay.row), dtype=torch.uint8, device=dev) if self.world > 1 else send
        with self.local._replay_lock:
            u = torch.rand(b_loc, device=dev, generator=smp._rng, dtype=smp._sum_tree._dtype)
            idx, _, leaf, pp = be.per_sample(smp._sum_tree.values, smp._min_tree.values, smp._max_capacity,
                                             smp._sum_tree.capacity, length, u, smp._beta, smp._semantics == "cpu",
                                             status=smp._status, want_aux=True)
            # with `peers` the rows are written into every rank's receive buffer by this very launch
            be.gather(st._leaves, idx, length, out=lay.leaf_views(send), peer_delta=peers)
            be.shard_pack(send, lay.meta, idx, leaf, pp, self.rank * self.shard_capacity, peer_delta=peers)
        self.local_index = idx
        self._send, self._recv, self._bs = send, recv, batch_size
        return send

    def exchange(self) -> torch.Tensor:
        """Close the exchange of sample().  nvlink: the rows are already on their way into every peer's buffer; a
        signal-pad barrier (one small kernel, stream-ordered) waits until everybody's have landed.  nccl: the ONE
        collective, an all-gather of the packed local draws."""
        if self.world > 1:
            if self._symm is not None and self._recv.data_ptr() == self._symm[0][self._parity].data_ptr():
                self._symm[1].barrier(channel=self._parity)
                self._parity ^= 1  # the next draw fills the other buffer; this one stays valid meanwhile
            else:
                self._dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        return self._recv

    def _use_nvlink(self, dev) -> bool:
        if self.world == 1 or dev.type != "cuda" or self.transport == "nccl":
            return False
        if self._symm is False:
            return False
        return True

    def _symmetric_buffers(self, batch_size: int, dev):
        """[2, B, row] symmetric receive buffer + the byte offsets from MY buffer to every rank's (0 for myself)."""
        if self._symm is not None and self._symm is not False and self._symm[0].shape[1] == batch_size:
            return self._symm
        try:
            import torch.distributed._symmetric_memory as symm_mem

            buf = symm_mem.empty((2, batch_size, self._layout.row), dtype=torch.uint8, device=dev)
            hdl = symm_mem.rendezvous(buf, group=self.group if self.group is not None else self._dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            peers = [p - ptrs[self.rank] for p in ptrs]
            if any(d % 16 for d in peers):
                raise RuntimeError("symmetric buffers are not 16-byte aligned relative to each other")
            self._symm = (buf, hdl, peers)
            self._parity = 0
        except Exception:
            if self.transport == "nvlink":
                raise
            self._symm = False  # fall back to the NCCL all-gather for good
            raise _NoSymmetricMemory()
        return self._symm


    def finalize(self):
        """Views of the gathered buffer as the global batch + importance weights over the whole sharded buffer."""
        lay, recv, smp = self._layout, self._recv, self.local.sampler
        leaves = lay.leaf_views(recv)
        batch = unflatten_data(leaves, self.local.storage._spec, (self._bs,))
        weight, gidx = ops.backend().shard_weights(recv, lay.meta, smp._beta)  # identical on every rank
        self._last_gidx = gidx
        if is_tensor_collection(batch):
            batch.set("index", gidx)
            batch.set("priority_weight", weight)
            return batch
        return batch, {"index": gidx, "priority_weight": weight}

    def sample(self, batch_size: int | None = None):
        self.local_draw(batch_size)
        self.exchange()
        return self.finalize()

    # ---- priority write-back ---------------------------------------------------------------------------
    def update_priority(self, index: torch.Tensor, priority) -> None:
        """``index`` holds GLOBAL indices; entries owned by other ranks are skipped inside the kernel."""
        index = torch.as_tensor(index, dtype=torch.long, device=self.device)
        if index is self._last_gidx and self._bs is not None:
            # the index vector of the batch sample() just returned: rows [rank*B/W, (rank+1)*B/W) are exactly the
            # draws from this shard, every other row belongs to another rank -- no need to scan them
            b_loc = self._bs // self.world
            lo = self.rank * b_loc
            priority = torch.as_tensor(priority, device=self.device)
            if priority.numel() > 1:
                priority = priority.reshape(-1)[lo:lo + b_loc]
            index = index[lo:lo + b_loc]
        self.local.sampler.update_priority(index, priority, storage=self.local.storage,
                                           index_base=self.rank * self.shard_capacity,
                                           index_limit=self.shard_capacity)

    def update_tensordict_priority(self, data) -> None:
        priority = data.get(self.local.priority_key)
        if priority.ndim > 1:
            priority = priority.reshape(priority.shape[0], -1).max(dim=1)[0]
        self.update_priority(data.get("index"), priority)
