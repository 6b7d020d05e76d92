"""Torch-tensor level entry to the C ABI of librlb200.so (include/rlb200.h).

Everything the host-side mirror of the TorchRL interface (rl_b200.data, rl_b200.objectives) computes
on the hot path goes through the functions below, which hand raw device pointers and the current
CUDA stream to the hand-written sm_100a kernels.  There is NO CPU implementation behind them: if the
shared library is missing, or a tensor is not on a CUDA device, they raise.

The only indirection is ``set_backend`` -- used by the CPU test-suite to plug in an oracle-backed
emulator (tests/_emul.py) so that the *host logic* (cursors, lengths, key plumbing, error behaviour)
can be exercised without a GPU.  The product never installs a backend other than ``CudaBackend``.
"""
from __future__ import annotations

import ctypes
from pathlib import Path
from typing import Sequence

import torch

_PKG = Path(__file__).resolve().parent
_SO = _PKG / "librlb200.so"

RLB_F32, RLB_F64 = 0, 1
GATHER_AUTO, GATHER_VECTOR, GATHER_BULK = 0, 1, 2
MAX_LEAVES = 24
STATUS_INDEX_OOB, STATUS_NONPOS_PSUM, STATUS_NONPOS_PMIN, STATUS_BACKOFF_FAIL = 1, 2, 4, 8
STATUS_EXCHANGE_TIMEOUT = 16
STATUS_FRAME_EVICTED = 32
FRAME_ENV_SHIFT = 40

_vp, _i64, _i32, _f64, _sz, _u32 = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
                                    ctypes.c_size_t, ctypes.c_uint32)

_SIGNATURES = {
    "rlb_version": (ctypes.c_int, []),
    "rlb_last_error": (ctypes.c_char_p, []),
    "rlb_device_sm_count": (ctypes.c_int, []),
    "rlb_l2_persist": (_i32, [_vp, _sz, _vp]),
    "rlb_tree_capacity": (_i64, [_i64]),
    "rlb_tree_update_workspace_bytes": (_sz, [_i64]),
    "rlb_tree_fill": (_i32, [_vp, _i64, _i32, _i32, _vp]),
    "rlb_tree_rebuild": (_i32, [_vp, _i64, _i32, _i32, _vp]),
    "rlb_tree_update": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp, _sz, _u32, _vp]),
    "rlb_tree_query": (_i32, [_vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _i32, _vp]),
    "rlb_tree_at": (_i32, [_vp, _i64, _i32, _vp, _vp, _i64, _vp]),
    "rlb_tree_scan_lower_bound": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _i64, _vp]),
    "rlb_per_sample": (_i32, [_vp, _vp, _i64, _i64, _i32, _i64, _vp, _i64, _f64, _i32, _vp, _vp, _vp, _vp,
                               _vp, _vp]),
    "rlb_per_update": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _f64, _f64, _vp, _vp, _vp, _sz, _u32, _i64, _i64,
                               _vp]),
    "rlb_shard_pack": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _vp, _i32, _vp]),
    "rlb_shard_weights": (_i32, [_vp, _i64, _i64, _i64, _f64, _vp, _vp, _vp, _vp, _i32, _f64, _vp, _vp]),
    "rlb_gather": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _i64, _i32, _vp, _vp]),
    "rlb_scatter": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _i64, _i64, _vp, _vp]),
    "rlb_gather_ex": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "rlb_framestack_push": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32,
                                   _i64, _i64, _vp]),
    "rlb_gae": (_i32, [_vp, _vp, _vp, _vp, _vp, _f64, _f64, _i64, _i64, _i64, _i32, _vp, _vp, _vp]),
    "rlb_td_lambda_return": (_i32, [_vp, _vp, _vp, _vp, _f64, _f64, _f64, _i64, _i64, _i64, _i32, _vp, _vp]),
    "rlb_affine_scan": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp]),
    "rlb_traj_table_workspace_bytes": (_sz, [_i64]),
    "rlb_traj_table": (_i32, [_vp, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rlb_slice_index": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp,
                               _vp, _vp, _vp, _vp, _vp]),
    "rlb_slice_mask_starts": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _i64, _vp]),
    "rlb_tree_update_range": (_i32, [_vp, _vp, _i64, _i32, _i64, _i64, _i64, _i32, _vp, _f64, _f64, _f64, _i32, _vp,
                                     _vp, _vp]),
    "rlb_extend": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _i64, _i32, _i32, _vp, _f64,
                          _f64, _f64, _i32, _vp, _vp, _vp]),
}


class FrameLeaf(ctypes.Structure):
    """``rlb_frame_leaf`` (include/rlb200.h)."""
    _fields_ = [("fpos", _vp), ("head", _vp), ("ring", _i64), ("offset", _i32), ("reserved", _i32)]


class GatherOpts(ctypes.Structure):
    """``rlb_gather_opts`` (include/rlb200.h)."""
    _fields_ = [("frames", _vp), ("peer_delta", _vp), ("multicast_delta", _i64), ("n_peers", _i32), ("reserved", _i32)]


def exported_symbols() -> list[str]:
    """Every symbol include/rlb200.h declares (checked against the built library in the CPU tests)."""
    return sorted(_SIGNATURES)


def load_library() -> ctypes.CDLL:
    """dlopen the in-tree librlb200.so and declare the prototypes.  Raises if it was never built."""
    if not _SO.exists():
        raise RuntimeError(
            f"{_SO} is missing: the B200 CUDA extension was not built. Run `python -c 'import "
            "__graft_entry__ as g; g.build()'` (needs nvcc). rl_b200 has no CPU fallback.")
    L = ctypes.CDLL(str(_SO))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


def _dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return RLB_F32
    if dt == torch.float64:
        return RLB_F64
    raise NotImplementedError(f"dtype {dt} not supported (fp32 / fp64 only)")


class CudaBackend:
    """ctypes binding of librlb200.so; every tensor must live on a CUDA device."""

    name = "cuda"

    def __init__(self):
        if not torch.cuda.is_available():
            raise RuntimeError("rl_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
        self.L = load_library()

    # -- helpers -----------------------------------------------------------------------------------
    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self.L.rlb_last_error().decode(errors="replace")
            raise RuntimeError(f"{what} failed with code {rc}: {msg}")

    @staticmethod
    def _cuda(*tensors: torch.Tensor) -> torch.device:
        dev = None
        for t in tensors:
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError(
                    f"rl_b200: expected a CUDA tensor, got device={t.device}; this engine has no CPU path.")
            if dev is None:
                dev = t.device
            elif t.device != dev:
                raise RuntimeError(f"rl_b200: tensors on different devices ({dev} vs {t.device})")
        return dev

    @staticmethod
    def _stream(dev: torch.device) -> int:
        """Raw handle of torch's current stream on `dev` (the binding Triton and the inductor runtime use: ~0.3 us, where
        ``torch.cuda.current_stream(dev).cuda_stream`` builds a Stream object first: ~7 us of every eager call)."""
        raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        if raw is None:
            return torch.cuda.current_stream(dev).cuda_stream
        return raw(dev.index if dev.index is not None else torch.cuda.current_device())

    @staticmethod
    def _p(t: torch.Tensor | None) -> int | None:
        return None if t is None else t.data_ptr()

    class _Guard:
        """Make `dev` the current CUDA device for the launch if it is not already."""

        def __init__(self, dev):
            self.ctx = None
            if dev.index is not None and dev.index != torch.cuda.current_device():
                self.ctx = torch.cuda.device(dev)

        def __enter__(self):
            if self.ctx is not None:
                self.ctx.__enter__()

        def __exit__(self, *a):
            if self.ctx is not None:
                self.ctx.__exit__(*a)

    def l2_persist(self, tensor: torch.Tensor | None, stream: "torch.cuda.Stream | None" = None) -> int:
        """Ask for `tensor`'s bytes to stay L2-resident for kernels launched on `stream` (default: current)."""
        if tensor is None:
            s = (stream or torch.cuda.current_stream()).cuda_stream
            return int(self.L.rlb_l2_persist(None, 0, s))
        dev = self._cuda(tensor)
        s = (stream or torch.cuda.current_stream(dev)).cuda_stream
        with self._Guard(dev):
            rc = int(self.L.rlb_l2_persist(tensor.data_ptr(), tensor.numel() * tensor.element_size(), s))
        if rc < 0:
            self._check(rc, "rlb_l2_persist")
        return rc

    # -- segment tree ------------------------------------------------------------------------------
    def tree_capacity(self, size: int) -> int:
        return int(self.L.rlb_tree_capacity(int(size)))

    def tree_new(self, size: int, is_min: bool, dtype: torch.dtype, device, out: torch.Tensor | None = None) -> torch.Tensor:
        cap = self.tree_capacity(size)
        tree = torch.empty(2 * cap, dtype=dtype, device=device) if out is None else out
        if tree.numel() != 2 * cap or tree.dtype != dtype or not tree.is_contiguous():
            raise RuntimeError("tree_new: `out` must be a contiguous tensor of 2*capacity elements of the tree dtype")
        dev = self._cuda(tree)
        with self._Guard(dev):
            self._check(self.L.rlb_tree_fill(tree.data_ptr(), cap, int(is_min), _dtype_code(dtype),
                                             self._stream(dev)), "rlb_tree_fill")
        return tree

    def tree_workspace(self, size: int, device) -> torch.Tensor:
        nbytes = int(self.L.rlb_tree_update_workspace_bytes(int(size)))
        return torch.zeros(nbytes // 8, dtype=torch.int64, device=device)

    def tree_rebuild(self, tree: torch.Tensor, capacity: int, is_min: bool) -> None:
        dev = self._cuda(tree)
        with self._Guard(dev):
            self._check(self.L.rlb_tree_rebuild(tree.data_ptr(), capacity, int(is_min), _dtype_code(tree.dtype),
                                                self._stream(dev)), "rlb_tree_rebuild")

    def tree_update(self, sum_tree, min_tree, capacity: int, index: torch.Tensor, value: torch.Tensor,
                    workspace: torch.Tensor | None, epoch: int) -> None:
        ref = sum_tree if sum_tree is not None else min_tree
        dev = self._cuda(sum_tree, min_tree, index, value, workspace)
        if value.dtype != ref.dtype:
            raise RuntimeError("value dtype must match the tree dtype")  # cuda_segment_tree.cu:163-164
        if index.dtype != torch.int64:
            raise RuntimeError("index must be an int64 tensor")  # cuda_segment_tree.cu:161-162
        n = index.numel()
        scalar = int(value.numel() == 1)
        if not scalar and value.numel() != n:
            raise RuntimeError("value must have one element or as many elements as index")
        index, value = index.contiguous(), value.contiguous()
        with self._Guard(dev):
            self._check(self.L.rlb_tree_update(
                self._p(sum_tree), self._p(min_tree), capacity, index.data_ptr(), value.data_ptr(), n, scalar,
                _dtype_code(ref.dtype), self._p(workspace), 0 if workspace is None else workspace.numel() * 8,
                epoch & 0xFFFFFFFF, self._stream(dev)), "rlb_tree_update")

    def tree_query(self, tree, size: int, capacity: int, is_min: bool, l: torch.Tensor, r: torch.Tensor,
                   root_fast_path: bool) -> torch.Tensor:
        dev = self._cuda(tree, l, r)
        if l.dtype != torch.int64 or r.dtype != torch.int64:
            raise RuntimeError("l and r must be int64 tensors")  # cuda_segment_tree.cu:176-178
        lc, rc_ = l.contiguous(), r.contiguous()
        out = torch.empty(lc.shape, dtype=tree.dtype, device=dev)
        with self._Guard(dev):
            self._check(self.L.rlb_tree_query(tree.data_ptr(), size, capacity, int(is_min), _dtype_code(tree.dtype),
                                              lc.data_ptr(), rc_.data_ptr(), out.data_ptr(), lc.numel(),
                                              int(root_fast_path), self._stream(dev)), "rlb_tree_query")
        return out

    def tree_at(self, tree, capacity: int, index: torch.Tensor) -> torch.Tensor:
        dev = self._cuda(tree, index)
        ic = index.contiguous()
        out = torch.empty(ic.shape, dtype=tree.dtype, device=dev)
        with self._Guard(dev):
            self._check(self.L.rlb_tree_at(tree.data_ptr(), capacity, _dtype_code(tree.dtype), ic.data_ptr(),
                                           out.data_ptr(), ic.numel(), self._stream(dev)), "rlb_tree_at")
        return out

    def tree_scan_lower_bound(self, tree, size: int, capacity: int, value: torch.Tensor) -> torch.Tensor:
        dev = self._cuda(tree, value)
        if value.dtype != tree.dtype:
            raise RuntimeError("value dtype must match the tree dtype")  # cuda_segment_tree.cu:190-192
        vc = value.contiguous()
        out = torch.empty(vc.shape, dtype=torch.int64, device=dev)
        with self._Guard(dev):
            self._check(self.L.rlb_tree_scan_lower_bound(tree.data_ptr(), size, capacity, _dtype_code(tree.dtype),
                                                         vc.data_ptr(), out.data_ptr(), vc.numel(),
                                                         self._stream(dev)), "rlb_tree_scan_lower_bound")
        return out

    # -- fused sampler arithmetic ------------------------------------------------------------------
    def per_sample(self, sum_tree, min_tree, size: int, capacity: int, length: int, u: torch.Tensor, beta: float,
                   cpu_semantics: bool, status: torch.Tensor | None = None, want_aux: bool = False,
                   out: tuple | None = None):
        """``out=(index, weight, leaf, psum_pmin)``: caller-owned result tensors (persistent buffers of a captured or
        multi-stream step)."""
        dev = self._cuda(sum_tree, min_tree, u, status)
        B = u.numel()
        u = u.contiguous()
        if out is not None:
            index, weight, leaf, pp = out
            want_aux = True
        else:
            index = torch.empty(B, dtype=torch.int64, device=dev)
            weight = torch.empty(B, dtype=torch.float32, device=dev)
            leaf = torch.empty(B, dtype=sum_tree.dtype, device=dev) if want_aux else None
            pp = torch.empty(2, dtype=sum_tree.dtype, device=dev) if want_aux else None
        with self._Guard(dev):
            self._check(self.L.rlb_per_sample(
                sum_tree.data_ptr(), min_tree.data_ptr(), size, capacity, _dtype_code(sum_tree.dtype), length,
                u.data_ptr(), B, float(beta), int(cpu_semantics), index.data_ptr(), weight.data_ptr(),
                self._p(leaf), self._p(pp), self._p(status), self._stream(dev)), "rlb_per_sample")
        if want_aux:
            return index, weight, leaf, pp
        return index, weight

    def per_update(self, sum_tree, min_tree, capacity: int, index: torch.Tensor, priority: torch.Tensor,
                   alpha: float, eps: float, max_out: torch.Tensor | None, workspace, epoch: int,
                   index_base: int = 0, index_limit: int = -1) -> None:
        dev = self._cuda(sum_tree, min_tree, index, priority, max_out, workspace)
        if priority.dtype != torch.float32 or (sum_tree is not None and sum_tree.dtype != torch.float32):
            raise NotImplementedError("fused per_update is fp32 only")
        n = index.numel()
        scalar = int(priority.numel() == 1)
        index, priority = index.contiguous(), priority.contiguous()
        scratch = torch.empty(n, dtype=torch.float32, device=dev) if n > 8192 else None
        with self._Guard(dev):
            self._check(self.L.rlb_per_update(
                self._p(sum_tree), self._p(min_tree), capacity, index.data_ptr(), priority.data_ptr(), n, scalar,
                float(alpha), float(eps), self._p(scratch), self._p(max_out), self._p(workspace),
                0 if workspace is None else workspace.numel() * 8, epoch & 0xFFFFFFFF, int(index_base),
                int(index_limit), self._stream(dev)), "rlb_per_update")

    # -- sharded minibatch trailer -----------------------------------------------------------------
    def shard_pack(self, rows: torch.Tensor, meta_offset: int, index, leaf, psum_pmin, index_base: int,
                   peer_delta: Sequence[int] | None = None, flags: torch.Tensor | None = None,
                   seq_counter: torch.Tensor | None = None, rank: int = 0) -> None:
        """Trailers of the local rows (into every peer with ``peer_delta``) and, with ``flags``, the release of this
        draw's sequence number into every peer's flag array (the split-phase exchange's "my rows are there")."""
        dev = self._cuda(rows, index, leaf, psum_pmin, flags, seq_counter)
        n_peers = 0 if peer_delta is None else len(peer_delta)
        peers = (ctypes.c_int64 * n_peers)(*peer_delta) if n_peers else None
        with self._Guard(dev):
            self._check(self.L.rlb_shard_pack(rows.data_ptr(), rows.stride(0), meta_offset, index.data_ptr(),
                                              leaf.data_ptr(), psum_pmin.data_ptr(), int(index_base), rows.shape[0],
                                              peers, n_peers, self._p(flags), self._p(seq_counter), int(rank),
                                              self._stream(dev)), "rlb_shard_pack")

    def shard_weights(self, rows: torch.Tensor, meta_offset: int, beta: float, flags: torch.Tensor | None = None,
                      wait_counter: torch.Tensor | None = None, n_ranks: int = 0, timeout_s: float = 10.0,
                      status: torch.Tensor | None = None, out: tuple | None = None):
        """Importance weights + global indices of the gathered rows; with ``flags`` the kernel first waits (acquire)
        until every rank has published the draw it finalises."""
        dev = self._cuda(rows, flags, wait_counter, status)
        B = rows.shape[0]
        if out is None:
            weight = torch.empty(B, dtype=torch.float32, device=dev)
            gidx = torch.empty(B, dtype=torch.int64, device=dev)
        else:
            weight, gidx = out
        with self._Guard(dev):
            self._check(self.L.rlb_shard_weights(rows.data_ptr(), rows.stride(0), meta_offset, B, float(beta),
                                                 weight.data_ptr(), gidx.data_ptr(), self._p(flags),
                                                 self._p(wait_counter), int(n_ranks), float(timeout_s),
                                                 self._p(status), self._stream(dev)),
                        "rlb_shard_weights")
        return weight, gidx

    # -- storage rows ------------------------------------------------------------------------------
    def _rows(self, who, big: Sequence[torch.Tensor], small: Sequence[torch.Tensor], index: torch.Tensor,
              length: int, status, mode: int = GATHER_AUTO):
        """big[k] is the [N, ...] storage leaf, small[k] the [B, ...] batch leaf (rows may be strided)."""
        dev = self._cuda(index, status, *big, *small)
        if index.dtype != torch.int64:
            raise RuntimeError("index must be an int64 tensor")
        index = index.contiguous()
        B = index.numel()
        with self._Guard(dev):
            for lo in range(0, len(big), MAX_LEAVES):
                bs, ss = big[lo:lo + MAX_LEAVES], small[lo:lo + MAX_LEAVES]
                n = len(bs)
                P, I = ctypes.c_void_p * n, ctypes.c_int64 * n
                rowb = I(*[b.element_size() * (b[0].numel() if b.ndim > 1 else 1) for b in bs])
                bstride = I(*[b.stride(0) * b.element_size() for b in bs])
                sstride = I(*[(s.stride(0) if s.ndim > 0 and s.shape[0] > 1 else (s[0].numel() if s.ndim > 1 else 1))
                              * s.element_size() for s in ss])
                bigp, smallp = P(*[b.data_ptr() for b in bs]), P(*[s.data_ptr() for s in ss])
                if who == "rlb_gather":
                    rc = self.L.rlb_gather(bigp, smallp, rowb, bstride, sstride, None, 0, n, index.data_ptr(), B, length,
                                           mode, self._p(status), self._stream(dev))
                else:
                    rc = self.L.rlb_scatter(smallp, bigp, rowb, bstride, n, index.data_ptr(), B, length,
                                            self._p(status), self._stream(dev))
                self._check(rc, who)

    @staticmethod
    def _check_rows(t: torch.Tensor) -> None:
        if t.ndim < 1 or (t.ndim > 1 and t.shape[0] > 0 and not t.is_contiguous() and not t[0].is_contiguous()):
            raise RuntimeError("leaves must be [N, ...] tensors whose rows are contiguous")

    def gather(self, leaves: Sequence[torch.Tensor], index: torch.Tensor, length: int, mode: int = GATHER_AUTO,
               status: torch.Tensor | None = None, out: Sequence[torch.Tensor] | None = None,
               peer_delta: Sequence[int] | None = None, multicast_delta: int = 0) -> list[torch.Tensor]:
        """out[k][b] = leaves[k][index[b]].  `out` may be given (rows may be strided views into a packed buffer);
        `peer_delta` (byte offsets, including 0) replicates every written byte into NVLink peer buffers;
        `multicast_delta`: offset of the NVLink multicast alias of `out` (wide rows are then stored once)."""
        return self.gather_plan(leaves).run(index, length, mode=mode, status=status, out=out, peer_delta=peer_delta,
                                            multicast_delta=multicast_delta)

    def gather_plan(self, leaves: Sequence[torch.Tensor], frames: Sequence | None = None) -> "GatherPlan":
        """Pre-marshalled source side of rlb_gather for a fixed set of storage leaves (pointers, row sizes and
        strides do not change between samples); storages cache it.  ``frames[k] = (fpos, head, ring, offset)`` marks
        leaf k as a frame of the de-duplicated frame-stack storage (leaves[k] is the frame pool), ``None`` otherwise."""
        return GatherPlan(self, leaves, frames)

    def framestack_push(self, obs: torch.Tensor, next_obs: torch.Tensor, is_init, done, last_done: torch.Tensor,
                        head: torch.Tensor, pool: torch.Tensor, n_envs: int, layout: int, k: int, ring: int) -> torch.Tensor:
        """Logs the frames of ``obs.shape[0]`` transitions into ``pool`` and returns their frame words (int64 [n])."""
        dev = self._cuda(obs, next_obs, is_init, done, last_done, head, pool)
        n = obs.shape[0]
        fpos = torch.empty(n, dtype=torch.int64, device=dev)
        if n == 0:
            return fpos
        scratch = torch.empty(n, dtype=torch.uint8, device=dev)
        for t in (obs, next_obs):
            if not t[0].is_contiguous():
                raise RuntimeError("framestack_push: the frame stacks of a transition must be contiguous")
        frame_bytes = pool[0].numel() * pool.element_size()
        with self._Guard(dev):
            self._check(self.L.rlb_framestack_push(
                obs.data_ptr(), next_obs.data_ptr(), obs.stride(0) * obs.element_size(),
                next_obs.stride(0) * next_obs.element_size(), self._p(is_init), self._p(done), last_done.data_ptr(),
                head.data_ptr(), pool.data_ptr(), fpos.data_ptr(), scratch.data_ptr(), n, n_envs, layout, k, frame_bytes,
                ring, self._stream(dev)), "rlb_framestack_push")
        return fpos

    def scatter(self, leaves: Sequence[torch.Tensor], data: Sequence[torch.Tensor], index: torch.Tensor, length: int,
                status: torch.Tensor | None = None) -> None:
        if index.numel():
            pairs = [(t, d.contiguous()) for t, d in zip(leaves, data) if t.numel() > 0 and d.numel() > 0]
            for t, _ in pairs:
                self._check_rows(t)
            if pairs:
                self._rows("rlb_scatter", [t for t, _ in pairs], [d for _, d in pairs], index, length, status)

    # -- write path --------------------------------------------------------------------------------
    def tree_update_range(self, rng: "RangeUpdate", start: int, n: int, modulo: int) -> None:
        """Priority write of the slots (start + arange(n)) % modulo (``rlb_tree_update_range``)."""
        dev = self._cuda(rng.sum, rng.mn, rng.value, rng.max_buf, rng.ticket)
        with self._Guard(dev):
            self._check(self.L.rlb_tree_update_range(
                self._p(rng.sum), self._p(rng.mn), rng.capacity, rng.dtype_code, start, n, modulo, rng.mode,
                self._p(rng.value), rng.alpha, rng.eps, rng.first_default, int(rng.has_max), self._p(rng.max_buf),
                self._p(rng.ticket), self._stream(dev)), "rlb_tree_update_range")

    def extend(self, stores: Sequence[torch.Tensor], data: Sequence[torch.Tensor], cursor: int, n: int,
               max_size: int, rng: "RangeUpdate | None" = None) -> None:
        """stores[k][(cursor + b) % max_size] = data[k][b] for every leaf and, with ``rng``, the priority write of the
        same slots -- one launch (``rlb_extend``; more than RLB_MAX_LEAVES leaves take one launch per group)."""
        pairs = [(t, d) for t, d in zip(stores, data) if t.numel() > 0 and d.numel() > 0]
        for t, d in pairs:
            self._check_rows(t)
            self._check_rows(d)
        extra = () if rng is None else (rng.sum, rng.mn, rng.value, rng.max_buf, rng.ticket)
        dev = self._cuda(*extra, *[t for t, _ in pairs], *[d for _, d in pairs])
        with self._Guard(dev):
            groups = [pairs[lo:lo + MAX_LEAVES] for lo in range(0, len(pairs), MAX_LEAVES)] or [[]]
            for gi, grp in enumerate(groups):
                k = len(grp)
                P, I = ctypes.c_void_p * max(k, 1), ctypes.c_int64 * max(k, 1)
                rowb = I(*[t.element_size() * (t[0].numel() if t.ndim > 1 else 1) for t, _ in grp])
                dstride = I(*[t.stride(0) * t.element_size() for t, _ in grp])
                sstride = I(*[(d.stride(0) if d.shape[0] > 1 else (d[0].numel() if d.ndim > 1 else 1))
                              * d.element_size() for _, d in grp])
                r = rng if gi == 0 else None     # the trees ride with the first group
                self._check(self.L.rlb_extend(
                    P(*[d.data_ptr() for _, d in grp]), P(*[t.data_ptr() for t, _ in grp]), rowb, dstride, sstride,
                    k, cursor, n, max_size,
                    self._p(r.sum) if r else None, self._p(r.mn) if r else None, r.capacity if r else 0,
                    r.dtype_code if r else 0, r.mode if r else 0, self._p(r.value) if r else None,
                    r.alpha if r else 0.0, r.eps if r else 0.0, r.first_default if r else 0.0,
                    int(r.has_max) if r else 0, self._p(r.max_buf) if r else None,
                    self._p(r.ticket) if r else None, self._stream(dev)), "rlb_extend")

    # -- trajectory slices -------------------------------------------------------------------------
    def traj_workspace(self, L: int, device) -> torch.Tensor:
        n = int(self.L.rlb_traj_table_workspace_bytes(int(L)))
        return torch.zeros((n + 7) // 8, dtype=torch.int64, device=device)

    def traj_table(self, signal: torch.Tensor, by_id: bool, L: int, at_capacity: bool, cursor: int, min_len: int,
                   keep_long_only: bool, table: torch.Tensor, counts: torch.Tensor, workspace: torch.Tensor) -> None:
        """table: int64 [3, >= L] (start, stop, length rows), counts: int64 [2]; see rlb_traj_table."""
        dev = self._cuda(signal, table, counts, workspace)
        if by_id:
            if signal.dtype != torch.int64:
                signal = signal.to(torch.int64)
        elif signal.dtype != torch.uint8:
            signal = signal.view(torch.uint8) if signal.dtype == torch.bool else (signal != 0).view(torch.uint8)
        signal = signal.contiguous()
        if signal.numel() < L or table.shape[1] < L or not table.is_contiguous():
            raise RuntimeError("traj_table: signal / table shorter than the storage")
        with self._Guard(dev):
            self._check(self.L.rlb_traj_table(signal.data_ptr(), 1 if by_id else 0, L, int(at_capacity), cursor, min_len,
                                              int(keep_long_only), table[0].data_ptr(), table[1].data_ptr(),
                                              table[2].data_ptr(), counts.data_ptr(), workspace.data_ptr(),
                                              workspace.numel() * 8, self._stream(dev)), "rlb_traj_table")

    def slice_index(self, start, length, n_traj: int, traj_draw, u, seq_length: int, storage_length: int,
                    variable: bool = False, pad_output: bool = False, out_offset=None, total: int | None = None,
                    want_index: bool = True, flags: tuple | None = None, span: tuple = (0, 0)):
        """Returns (index int64[n], truncated bool[n, 1], mask bool[n] | None, seq int64[num_slices]); with
        ``flags=(done_leaf | None, terminated_leaf | None)`` (the storage's one-byte-per-slot flags) two more entries:
        ``done[index] | truncated`` and ``terminated[index]`` as bool [n, 1]."""
        dev = self._cuda(start, length, traj_draw, u, out_offset, *(f for f in (flags or ()) if f is not None))
        S = traj_draw.numel()
        seq = torch.empty(S, dtype=torch.int64, device=dev)
        index = trunc = mask = None
        if want_index:
            n = S * seq_length if (not variable or pad_output) else int(total)
            index = torch.empty(n, dtype=torch.int64, device=dev)
            trunc = torch.empty((n, 1), dtype=torch.bool, device=dev)
            mask = torch.empty(n, dtype=torch.bool, device=dev) if (variable and pad_output) else None
        done_src = term_src = done_out = term_out = None
        if flags is not None and want_index:
            done_src, term_src = (None if f is None else f.contiguous().view(torch.uint8) for f in flags)
            both = torch.empty((2, n, 1), dtype=torch.bool, device=dev)
            done_out, term_out = both[0], both[1]
        with self._Guard(dev):
            self._check(self.L.rlb_slice_index(start.data_ptr(), length.data_ptr(), n_traj, traj_draw.data_ptr(),
                                               u.data_ptr(), S, seq_length, storage_length, int(variable),
                                               int(pad_output), int(span[0]), int(span[1]), self._p(out_offset),
                                               self._p(index), self._p(trunc),
                                               self._p(mask), seq.data_ptr(), self._p(done_src), self._p(term_src),
                                               self._p(done_out), self._p(term_out), self._stream(dev)),
                        "rlb_slice_index")
        if flags is not None and want_index:
            return index, trunc, mask, seq, done_out, term_out
        return index, trunc, mask, seq

    def slice_mask_starts(self, masked_tree: torch.Tensor, capacity: int, stop, length, n_traj: int, seq_length: int,
                          ring_length: int) -> None:
        """Zero, in the LEAF level of ``masked_tree`` (a copy of a sum tree), the starts from which a slice of
        ``seq_length`` steps would leave its trajectory (``rlb_slice_mask_starts``)."""
        dev = self._cuda(masked_tree, stop, length)
        leaves = masked_tree[capacity:]
        with self._Guard(dev):
            self._check(self.L.rlb_slice_mask_starts(leaves.data_ptr(), _dtype_code(masked_tree.dtype), stop.data_ptr(),
                                                     length.data_ptr(), n_traj, seq_length, ring_length,
                                                     self._stream(dev)), "rlb_slice_mask_starts")

    # -- GAE ---------------------------------------------------------------------------------------
    def gae(self, v, nv, r, done, term, gamma: float, gammalmbda: float, rows: int, T: int, F: int):
        dev = self._cuda(v, nv, r, done, term)
        adv, tgt = torch.empty_like(v), torch.empty_like(v)
        with self._Guard(dev):
            self._check(self.L.rlb_gae(v.data_ptr(), nv.data_ptr(), r.data_ptr(), done.data_ptr(), term.data_ptr(),
                                       float(gamma), float(gammalmbda), rows, T, F, _dtype_code(v.dtype),
                                       adv.data_ptr(), tgt.data_ptr(), self._stream(dev)), "rlb_gae")
        return adv, tgt

    def td_lambda_return(self, nv, r, done, term, gamma: float, gammalmbda: float, one_minus_lmbda: float, rows: int,
                         T: int, F: int):
        dev = self._cuda(nv, r, done, term)
        ret = torch.empty_like(nv)
        with self._Guard(dev):
            self._check(self.L.rlb_td_lambda_return(nv.data_ptr(), r.data_ptr(), done.data_ptr(), term.data_ptr(),
                                                    float(gamma), float(gammalmbda), float(one_minus_lmbda), rows, T, F,
                                                    _dtype_code(nv.dtype), ret.data_ptr(), self._stream(dev)),
                        "rlb_td_lambda_return")
        return ret

    def affine_scan(self, d, c, rows: int, T: int, F: int):
        dev = self._cuda(d, c)
        out = torch.empty_like(d)
        with self._Guard(dev):
            self._check(self.L.rlb_affine_scan(d.data_ptr(), c.data_ptr(), rows, T, F, _dtype_code(d.dtype),
                                               out.data_ptr(), self._stream(dev)), "rlb_affine_scan")
        return out


RANGE_VALUE, RANGE_PRIORITY, RANGE_DEFAULT = 0, 1, 2


class RangeUpdate:
    """Tree-side arguments of ``rlb_tree_update_range`` / ``rlb_extend`` (see include/rlb200.h)."""

    __slots__ = ("sum", "mn", "capacity", "dtype_code", "mode", "value", "alpha", "eps", "first_default", "has_max",
                 "max_buf", "ticket")

    def __init__(self, sum, mn, capacity, mode, value=None, alpha=1.0, eps=0.0, first_default=1.0, has_max=False,
                 max_buf=None, ticket=None):
        self.sum, self.mn, self.capacity, self.mode, self.value = sum, mn, int(capacity), int(mode), value
        self.dtype_code = _dtype_code((sum if sum is not None else mn).dtype)
        self.alpha, self.eps, self.first_default = float(alpha), float(eps), float(first_default)
        self.has_max, self.max_buf, self.ticket = bool(has_max), max_buf, ticket


class GatherPlan:
    """Source-side arguments of ``rlb_gather`` marshalled once for a set of [N, ...] leaves."""

    def __init__(self, be: CudaBackend, leaves: Sequence[torch.Tensor], frames: Sequence | None = None):
        self.be = be
        self.leaves = list(leaves)
        self.frames = None if frames is None or not any(f is not None for f in frames) else list(frames)
        for t in self.leaves:
            be._check_rows(t)
        self.dev = be._cuda(*self.leaves)
        self.tails = [tuple(t.shape[1:]) for t in self.leaves]
        self.dtypes = [t.dtype for t in self.leaves]
        self.chunks = []
        # leaves whose rows hold no bytes (a zero-sized feature dim) have nothing to move and no device pointer
        live = [k for k, t in enumerate(self.leaves) if t.numel() > 0]
        for lo in range(0, len(live), MAX_LEAVES):
            ks = live[lo:lo + MAX_LEAVES]
            ts = [self.leaves[k] for k in ks]
            n = len(ts)
            P, I = ctypes.c_void_p * n, ctypes.c_int64 * n
            rowb = [t.element_size() * (t[0].numel() if t.ndim > 1 else 1) for t in ts]
            fr = None
            if self.frames is not None and any(self.frames[k] is not None for k in ks):
                fr = (FrameLeaf * n)()
                for slot, k in enumerate(ks):
                    f = self.frames[k]
                    if f is not None:
                        fpos, head, ring, offset = f
                        fr[slot] = FrameLeaf(fpos.data_ptr(), None if head is None else head.data_ptr(), int(ring),
                                             int(offset), 0)
            self.chunks.append((ks, n, P, P(*[t.data_ptr() for t in ts]), I(*rowb),
                                I(*[t.stride(0) * t.element_size() for t in ts]), I, fr))

    def run(self, index: torch.Tensor, length: int, mode: int = GATHER_AUTO, status: torch.Tensor | None = None,
            out: Sequence[torch.Tensor] | None = None, peer_delta: Sequence[int] | None = None,
            multicast_delta: int = 0) -> list[torch.Tensor]:
        be = self.be
        n_peers = 0 if peer_delta is None else len(peer_delta)
        peers = (ctypes.c_int64 * n_peers)(*peer_delta) if n_peers else None
        if not index.is_cuda or index.device != self.dev:
            raise RuntimeError(f"rl_b200: index must live on {self.dev}, got {index.device}; there is no CPU path.")
        if index.dtype != torch.int64:
            raise RuntimeError("index must be an int64 tensor")
        if not index.is_contiguous():
            index = index.contiguous()
        B = index.numel()
        dev = self.dev
        strided = out is not None
        if out is None:
            out = [torch.empty((B, *tail), dtype=dt, device=dev) for tail, dt in zip(self.tails, self.dtypes)]
        else:
            for o, tail, dt in zip(out, self.tails, self.dtypes):
                be._check_rows(o)
                if tuple(o.shape) != (B, *tail) or o.dtype != dt or o.device != dev:
                    raise RuntimeError("gather: `out` leaf has the wrong shape, dtype or device")
        if B == 0:
            return list(out)
        if length <= 0:
            raise RuntimeError("rl_b200: cannot index an empty storage (len == 0)")
        stream = be._stream(dev)
        with be._Guard(dev):
            for ks, n, P, srcp, rowb, sstride, I, fr in self.chunks:
                outs = [out[k] for k in ks]
                dstp = P(*[o.data_ptr() for o in outs])
                dstride = None
                if strided:
                    dstride = I(*[(o.stride(0) if B > 1 else (o[0].numel() if o.ndim > 1 else 1)) * o.element_size()
                                  for o in outs])
                if fr is not None or multicast_delta:
                    opts = GatherOpts(None if fr is None else ctypes.cast(fr, _vp),
                                      None if peers is None else ctypes.cast(peers, _vp), int(multicast_delta), n_peers, 0)
                    be._check(be.L.rlb_gather_ex(srcp, dstp, rowb, sstride, dstride, n, ctypes.byref(opts),
                                                 index.data_ptr(), B, length, mode, be._p(status), stream),
                              "rlb_gather_ex")
                    continue
                be._check(be.L.rlb_gather(srcp, dstp, rowb, sstride, dstride, peers, n_peers, n, index.data_ptr(), B,
                                          length, mode, be._p(status), stream), "rlb_gather")
        return list(out)


_BACKEND = None


def backend():
    """The active backend; created on first use.  Fails loudly without the CUDA extension / a GPU."""
    global _BACKEND
    if _BACKEND is None:
        _BACKEND = CudaBackend()
    return _BACKEND


def set_backend(b) -> None:
    """TESTS ONLY: install an emulator backend (or None to restore the CUDA one)."""
    global _BACKEND
    _BACKEND = b
