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
    ranks are ONE symmetric allocation (``torch.distributed._symmetric_memory`` is used for the allocation and the
    peer addresses only), so every rank knows the address of its rows in every peer's buffer and the gather
    kernel's shared-memory stages are stored W times -- once locally, W-1 times through NVLink peer memory.
    Gather and all-gather are the SAME launch (``rlb_gather`` with ``peer_delta``).  The exchange is split-phase
    and closed by the library's own flags: ``rlb_shard_pack`` (trailers) ends with a system-scope release of the
    draw's sequence number into every peer's flag array, ``rlb_shard_weights`` (finalize) starts with an acquire
    spin on its own flags.  No NCCL call and no library barrier, so the whole step is capturable in one graph.
  * ``pipeline=True`` (split-phase across calls, like the reference's ``prefetch=1``): ``sample()`` issues draw k
    and returns the finalised batch k-1, so the NVLink transfer of a draw overlaps everything that runs until
    the next call and the flag wait never stalls.  Receive buffers rotate over ``n_buffers`` slots (4 when
    pipelined: a batch stays valid until the second next ``sample()``; 2 otherwise).
  * transport "nccl": ``torch.distributed.all_gather_into_tensor`` moves ``B/W * row`` bytes per rank (also the
    CPU/gloo test path).  Either way the leaves of the returned batch are strided views into the receive buffer.

``update_priority`` takes GLOBAL indices (replicated on every rank, or any subset): each rank rewrites the
ones it owns and skips the rest inside the kernel (negative local index) -- no collective, no sync.
``update_local_priority`` takes priorities for the rows this rank drew last (``local_index``): it only needs the
sampled indices, so it can be issued right after the tree kernel, concurrently with the exchange.
"""
from __future__ import annotations

import os

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
        transport ("auto" | "nvlink" | "nccl"): see the module docstring.
        pipeline (bool): ``sample()`` returns the PREVIOUS draw (the first call draws twice).
        n_buffers (int): receive-buffer slots of the nvlink transport (default 4 pipelined, 2 otherwise).
        exchange_timeout_s (float): bound of the in-kernel wait for the peers' rows.
        multicast (bool | "auto"): nvlink transport: store the wide rows ONCE through the NVLink-SHARP multicast mapping of
            the symmetric receive buffers (``multimem.st``; the switch replicates them into every rank) instead of one copy
            per rank -- NVLink egress per rank drops from (W - 1) x to 1 x the local draw.  "auto": when torch's symmetric
            memory exposes a multicast pointer AND there are at least 4 ranks (with 2 there is nothing to replicate and the
            DMA engine's unicast copy is faster); "probe": whenever the pointer exists; True: required (env
            ``RLB_SHARD_MULTICAST=0/1`` overrides).
        storage: this rank's shard storage (``shard_capacity`` slots); default ``LazyTensorStorage``.  With a
            ``FrameStackStorage`` the exchanged row carries the k + 1 distinct frames of a transition instead of both
            k-frame stacks (35.3 KB instead of 56.4 KB per Atari transition), and ``obs`` / ``next`` of the returned batch
            are two views of that window.
    """

    def __init__(self, *, alpha: float, beta: float, capacity: int, eps: float = 1e-8, priority_key: str = "td_error",
                 batch_size: int | None = None, device="cuda", generator=None, process_group=None,
                 transport: str = "auto", pipeline: bool = False, n_buffers: int | None = None,
                 exchange_timeout_s: float = 10.0, storage=None, multicast: bool | str = "auto"):
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
            storage=storage if storage is not None else LazyTensorStorage(self.shard_capacity, device=self.device),
            batch_size=None if batch_size is None else batch_size // self.world, generator=generator)
        if storage is not None and storage.max_size != self.shard_capacity:
            raise ValueError(f"the shard storage must hold {self.shard_capacity} slots, got {storage.max_size}")
        if transport not in ("auto", "nvlink", "nccl"):
            raise ValueError("transport must be 'auto', 'nvlink' or 'nccl'")
        self.transport = transport
        self.pipeline = bool(pipeline)
        self.n_buffers = int(n_buffers) if n_buffers is not None else (4 if pipeline else 2)
        if self.n_buffers < (3 if pipeline else 2):
            raise ValueError("n_buffers must be >= 2 (>= 3 when pipelined)")
        self.exchange_timeout_s = float(exchange_timeout_s)
        env = os.environ.get("RLB_SHARD_MULTICAST")
        self.multicast = multicast if env is None else {"0": False, "1": True}.get(env, "auto")
        self._mc_delta = 0       # byte offset from my symmetric allocation to its NVLink multicast alias (0: unicast copies)
        self._last_gidx = None
        self._symm = None        # (buffers [n_buffers, B, row], flags u64[W], handle, peer byte offsets)
        self._draws = 0          # host-side count of issued draws: picks the receive slot
        self._ctr = None         # device counters: [0] draws published, [1] draws waited for
        self._xstatus = None     # device status word of the exchange (RLB_STATUS_EXCHANGE_TIMEOUT)
        self._layout = None
        self._static = None
        self._cur = None         # the draw in flight: dict(send, recv, bs, nvlink)
        self._pending = None     # pipelined: the draw the NEXT sample() returns
        self._fin_stream = None
        self._x_stream = None    # the exchange (gather + NVLink broadcast + publish) runs here, beside the caller's stream
        self._x_done = {}        # slot -> event: that slot's previous exchange has left this GPU
        self._x_done_cap = {}    # the same for draws issued inside the CUDA-graph capture in progress
        self._res = {}           # slot -> persistent (index, weight, leaf, psum_pmin) of the draw using it
        #: nvlink transport: issue the exchange on an internal stream so that the caller's stream goes straight on to
        #: the priority write-back / the next draw (the trees do not depend on it).  ``join_exchange()`` joins it.
        self.overlap_exchange = True
        self.local_index = None  # local indices of this rank's last draw
        #: set to True to have ``local_draw`` record ``index_ready`` (a CUDA event) right after the tree kernel: work
        #: that only needs the sampled indices (``update_local_priority`` on a side stream) overlaps the exchange
        self.record_index_event = False
        self.index_ready = None

    # ---- writes: every rank feeds its own shard (data-parallel collectors) --------------------------------
    def extend(self, data) -> torch.Tensor:
        """Writes ``data`` into this rank's shard; returns the GLOBAL indices of the written slots."""
        self.join_exchange()   # an exchange still in flight reads the rows this write may replace
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
    # sample() = local_draw() -> exchange() -> finalize().  The stages are public so that a training step can place
    # them on streams of its choice (and, with the NCCL transport, issue the collective between two captured graphs).
    def _resolve_batch(self, batch_size):
        if batch_size is None:
            batch_size = self._batch_size
        if batch_size is None:
            raise RuntimeError("batch_size not specified.")
        if batch_size % self.world:
            raise ValueError(f"batch_size={batch_size} must be divisible by the world size {self.world}")
        return batch_size

    def local_draw(self, batch_size: int | None = None, *, static_buffers: bool = False,
                   slot: int | None = None) -> torch.Tensor:
        """Draw ``batch_size / world`` rows from this shard straight into the packed send buffer (returned); with the
        nvlink transport the same launch stores them into every peer's receive buffer and the trailer kernel
        publishes the draw.

        ``static_buffers=True`` reuses one send / receive buffer pair across calls (NCCL transport under CUDA-graph
        capture; the returned batch is then overwritten by the next sample).  ``slot`` picks the receive slot of the
        nvlink transport explicitly (captured steps: every rank must use the same slot for the same draw, and a
        slot must not come round again while its batch is still in use -- replay ``k * n_buffers`` captured steps
        cyclically); default: issued draws modulo ``n_buffers``.
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
            self._layout = _PackedLayout(st._packed_templates() if hasattr(st, "_packed_templates") else st._leaves)
        lay = self._layout
        peers = flags = None
        nv = None
        if self._use_nvlink(dev):
            try:
                nv = self._symmetric_buffers(batch_size, dev)
            except _NoSymmetricMemory:
                nv = None
        if nv is not None:
            bufs, flags, _, peers = nv
            recv = bufs[(self._draws if slot is None else slot) % self.n_buffers]
            send = recv[self.rank * b_loc:(self.rank + 1) * b_loc]  # my rows inside the gathered batch
        elif static_buffers:
            if self._static is None or self._static[0].shape[0] != b_loc:
                self._static = (torch.empty((b_loc, lay.row), dtype=torch.uint8, device=dev),
                                torch.empty((batch_size, lay.row), dtype=torch.uint8, device=dev))
            send, recv = self._static
        else:
            send = torch.empty((b_loc, lay.row), dtype=torch.uint8, device=dev)
            recv = torch.empty((batch_size, lay.row), dtype=torch.uint8, device=dev) if self.world > 1 else send
        slot_i = (self._draws if slot is None else slot) % self.n_buffers
        res = self._res.get(slot_i)
        if res is None or res[0].numel() != b_loc or res[0].device != dev:
            # persistent results of the tree kernel, one set per receive slot: they are read by the exchange on its own
            # stream and by write-backs on side streams, so they must not be caching-allocator temporaries
            dt = smp._sum_tree._dtype
            res = self._res[slot_i] = (torch.empty(b_loc, dtype=torch.int64, device=dev),
                                       torch.empty(b_loc, dtype=torch.float32, device=dev),
                                       torch.empty(b_loc, dtype=dt, device=dev), torch.empty(2, dtype=dt, device=dev))
        side = nv is not None and self.overlap_exchange and dev.type == "cuda"
        capturing = dev.type == "cuda" and torch.cuda.is_current_stream_capturing()
        with self.local._replay_lock:
            done = (self._x_done_cap if capturing else self._x_done).get(slot_i) if side else None
            if done is not None:
                # the exchange that used this slot's buffers last must have left the GPU before the tree kernel
                # overwrites them (inside a capture: only draws of the same capture are known, and join_exchange()
                # at its end orders everything else)
                torch.cuda.current_stream(dev).wait_event(done)
            u = torch.rand(b_loc, device=dev, generator=smp._rng, dtype=smp._sum_tree._dtype)
            idx, _, leaf, pp = be.per_sample(smp._sum_tree.values, smp._min_tree.values, smp._max_capacity,
                                             smp._sum_tree.capacity, length, u, smp._beta, smp._semantics == "cpu",
                                             status=smp._status, out=res)
            if self.record_index_event and dev.type == "cuda":
                if self.index_ready is None:
                    self.index_ready = torch.cuda.Event()
                self.index_ready.record(torch.cuda.current_stream(dev))

            def exchange_kernels():
                # with `peers` the rows are written into every rank's receive buffer by this very launch
                mc = self._mc_delta if peers is not None else 0
                if hasattr(st, "_gather_packed"):   # FrameStackStorage: k + 1 unique frames per transition on the wire
                    st._gather_packed(idx, lay.leaf_views(send), peer_delta=peers, multicast_delta=mc)
                else:
                    be.gather(st._leaves, idx, length, out=lay.leaf_views(send), peer_delta=peers, multicast_delta=mc)
                # trailers; with `flags`: "my rows of this draw are in your buffer" to every rank (release)
                be.shard_pack(send, lay.meta, idx, leaf, pp, self.rank * self.shard_capacity, peer_delta=peers,
                              flags=flags, seq_counter=None if flags is None else self._ctr[0:1], rank=self.rank)

            if side:
                if self._x_stream is None:
                    self._x_stream = torch.cuda.Stream(dev)
                cur = torch.cuda.current_stream(dev)
                self._x_stream.wait_stream(cur)
                with torch.cuda.stream(self._x_stream):
                    exchange_kernels()
                    table = self._x_done_cap if capturing else self._x_done
                    ev = table.get(slot_i)
                    if ev is None:
                        ev = table[slot_i] = torch.cuda.Event()
                    ev.record(self._x_stream)
            else:
                exchange_kernels()
        self.local_index = idx
        self._draws += 1
        self._cur = {"send": send, "recv": recv, "bs": batch_size, "nvlink": nv is not None, "exchanged": False}
        return send

    def join_exchange(self) -> None:
        """Make the current stream wait for every exchange issued so far (needed before the storage is written to, and
        before a CUDA-graph capture that issued draws ends).  A no-op when nothing runs beside the caller's stream."""
        if self._x_stream is not None:
            torch.cuda.current_stream(self._x_stream.device).wait_stream(self._x_stream)
        self._x_done_cap = {}

    def exchange(self) -> torch.Tensor:
        """nccl: the ONE collective, an all-gather of the packed local draws.  nvlink: nothing to issue -- the rows
        are already on their way into every peer's buffer; ``finalize`` waits for the flags."""
        cur = self._cur
        if self.world > 1 and not cur["nvlink"] and not cur["exchanged"]:
            self._dist.all_gather_into_tensor(cur["recv"], cur["send"], group=self.group)
        cur["exchanged"] = True
        return cur["recv"]

    def _use_nvlink(self, dev) -> bool:
        if self.world == 1 or dev.type != "cuda" or self.transport == "nccl":
            return False
        if self._symm is False:
            return False
        return True

    _FLAG_BYTES = 256  # one u64 per rank (RLB_MAX_PEERS = 16), padded so that the row buffers stay 128-B aligned

    def _symmetric_buffers(self, batch_size: int, dev):
        """Symmetric allocation  [flags u64[W] | n_buffers x [B, row]]  + the byte offsets from MY allocation to every
        rank's (0 for myself)."""
        if self._symm is not None and self._symm is not False and self._symm[0].shape[1] == batch_size:
            return self._symm
        try:
            import torch.distributed._symmetric_memory as symm_mem

            row = self._layout.row
            nbytes = self._FLAG_BYTES + self.n_buffers * batch_size * row
            raw = symm_mem.empty((nbytes,), dtype=torch.uint8, device=dev)
            hdl = symm_mem.rendezvous(raw, group=self.group if self.group is not None else self._dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            peers = [p - ptrs[self.rank] for p in ptrs]
            if any(d % 16 for d in peers) or raw.data_ptr() % 16:
                raise RuntimeError("symmetric buffers are not 16-byte aligned relative to each other")
            raw.zero_()
            torch.cuda.synchronize(dev)
            self._dist.barrier(group=self.group)   # nobody publishes into flags that are not zeroed yet
            mc = 0
            # "auto": from 4 ranks on -- with 2 the switch has nothing to replicate and the DMA engine's unicast copy is
            # faster than warp-issued multimem.st (measured: 25 us vs 44 us for the C2 exchange at N=2)
            if self.multicast in (True, "probe") or (self.multicast == "auto" and self.world >= 4):
                try:
                    mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                except Exception:  # a fabric / driver without NVLink SHARP: unicast copies
                    mc = 0
                if self.multicast is True and not mc:
                    raise RuntimeError("multicast=True, but the symmetric allocation has no multicast mapping")
            self._mc_delta = (mc - raw.data_ptr()) if mc else 0
            if self._mc_delta % 16:
                self._mc_delta = 0
            flags = raw[:self._FLAG_BYTES].view(torch.int64)[:self.world]
            bufs = raw[self._FLAG_BYTES:].view(self.n_buffers, batch_size, row)
            self._ctr = torch.zeros(2, dtype=torch.int64, device=dev)
            self._xstatus = torch.zeros(1, dtype=torch.int32, device=dev)
            self._symm = (bufs, flags, hdl, peers)
            self._raw = raw
        except Exception as err:
            if self.transport == "nvlink":
                raise
            import warnings

            warnings.warn(f"ShardedPrioritizedReplayBuffer: symmetric memory unavailable ({type(err).__name__}: {err}); "
                          "using the NCCL all-gather transport")
            self._symm = False  # fall back to the NCCL all-gather for good
            raise _NoSymmetricMemory()
        return self._symm

    def finalize(self, draw: dict | None = None):
        """Views of the gathered buffer as the global batch + importance weights over the whole sharded buffer.  With
        the nvlink transport the weights kernel first waits until every rank's rows of this draw have landed."""
        cur = self._cur if draw is None else draw
        lay, recv, smp = self._layout, cur["recv"], self.local.sampler
        leaves = lay.leaf_views(recv)
        st = self.local.storage
        batch = st._unpack(leaves, (cur["bs"],)) if hasattr(st, "_unpack") else unflatten_data(leaves, st._spec, (cur["bs"],))
        if cur["nvlink"]:
            weight, gidx = ops.backend().shard_weights(recv, lay.meta, smp._beta, flags=self._symm[1],
                                                       wait_counter=self._ctr[1:2], n_ranks=self.world,
                                                       timeout_s=self.exchange_timeout_s, status=self._xstatus,
                                                       out=cur.get("out"))
        else:
            weight, gidx = ops.backend().shard_weights(recv, lay.meta, smp._beta)  # identical on every rank
        self._last_gidx = gidx
        self._last_bs = cur["bs"]
        if is_tensor_collection(batch):
            batch.set("index", gidx)
            batch.set("priority_weight", weight)
            return batch
        return batch, {"index": gidx, "priority_weight": weight}

    def check_exchange(self) -> None:
        """Synchronise and raise if a peer's rows ever failed to arrive within ``exchange_timeout_s``."""
        if self._xstatus is not None and int(self._xstatus.item()) & ops.STATUS_EXCHANGE_TIMEOUT:
            self._xstatus.zero_()
            raise RuntimeError("sharded exchange: a peer's rows did not arrive in time")

    def sample_now(self, batch_size: int | None = None, *, slot: int | None = None):
        """Draw, exchange and finalise in one go (what ``sample`` does when not pipelined)."""
        if self._pending is not None:
            raise RuntimeError("a pipelined draw is pending: call sample() (or flush()) first")
        self.local_draw(batch_size, slot=slot)
        self.exchange()
        return self.finalize()

    def sample(self, batch_size: int | None = None, *, slot: int | None = None):
        """The global minibatch.  Pipelined: issues draw k and returns the finalised draw k-1 (the first call issues two
        draws); the finalisation runs on a side stream forked from / joined to the current one, so in a captured step it
        is not chained behind the new draw."""
        if not self.pipeline:
            return self.sample_now(batch_size, slot=slot)
        if self._pending is None:
            self.local_draw(batch_size, slot=None if slot is None else slot - 1)
            self.exchange()
            self._pending = self._cur
        prev = self._pending
        if slot is not None and prev["nvlink"]:
            # explicit slots (captured steps replayed cyclically): the draw this call finalises is the one the step
            # with slot - 1 issued, whatever Python call happened to precede this one at capture time
            prev = dict(prev, recv=self._symm[0][(slot - 1) % self.n_buffers])
        dev = self.local.sampler._sum_tree.device
        if dev.type == "cuda" and prev["nvlink"]:
            # outputs are allocated on the caller's stream, the wait + weights kernel runs beside the new draw
            prev["out"] = (torch.empty(prev["bs"], dtype=torch.float32, device=dev),
                           torch.empty(prev["bs"], dtype=torch.int64, device=dev))
            if self._fin_stream is None:
                self._fin_stream = torch.cuda.Stream(dev)
            main = torch.cuda.current_stream(dev)
            self._fin_stream.wait_stream(main)
            with torch.cuda.stream(self._fin_stream):
                out = self.finalize(prev)
            self.local_draw(batch_size, slot=slot)
            self.exchange()
            main.wait_stream(self._fin_stream)
        else:
            self.local_draw(batch_size, slot=slot)
            self.exchange()
            out = self.finalize(prev)
        self._pending = self._cur
        return out

    def flush(self):
        """Pipelined: finalise and return the pending draw without issuing a new one (``None`` if there is none)."""
        if self._pending is None:
            return None
        prev, self._pending = self._pending, None
        return self.finalize(prev)

    # ---- priority write-back ---------------------------------------------------------------------------
    def update_priority(self, index: torch.Tensor, priority) -> None:
        """``index`` holds GLOBAL indices; entries owned by other ranks are skipped inside the kernel."""
        index = torch.as_tensor(index, dtype=torch.long, device=self.device)
        last = self._last_gidx
        if (last is not None and index.numel() == last.numel() and index.data_ptr() == last.data_ptr()
                and index.is_contiguous()):
            # the index vector of the batch sample() returned last (the tensor itself or a view of it): rows
            # [rank*B/W, (rank+1)*B/W) are exactly the draws from this shard, every other row belongs to another
            # rank -- no need to scan them.  Anything else (a clone, a subset, a permutation) takes the general path.
            b_loc = self._last_bs // self.world
            lo = self.rank * b_loc
            priority = torch.as_tensor(priority, device=self.device)
            if priority.numel() > 1:
                priority = priority.reshape(-1)[lo:lo + b_loc]
            index = index.reshape(-1)[lo:lo + b_loc]
        self.local.sampler.update_priority(index, priority, storage=self.local.storage,
                                           index_base=self.rank * self.shard_capacity,
                                           index_limit=self.shard_capacity)

    def update_local_priority(self, priority) -> None:
        """Priorities for the rows THIS rank drew in its latest ``local_draw`` (``local_index``, in draw order).  Needs
        nothing but the sampled indices, so it may be issued as soon as ``index_ready`` has fired -- concurrently with
        the gather / exchange of the same draw."""
        if self.local_index is None:
            raise RuntimeError("no draw yet")
        self.local.sampler.update_priority(self.local_index, priority, storage=self.local.storage)

    def update_tensordict_priority(self, data) -> None:
        priority = data.get(self.local.priority_key)
        if priority.ndim > 1:
            priority = priority.reshape(priority.shape[0], -1).max(dim=1)[0]
        self.update_priority(data.get("index"), priority)
