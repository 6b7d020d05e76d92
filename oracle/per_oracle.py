"""TEST INFRASTRUCTURE -- prioritized-replay oracle (never imported by the product).

``OracleTree`` wraps the C restatement of the reference segment tree (oracle/rlb_oracle.c).
``OraclePrioritizedSampler`` restates the *Python glue* of the reference
``PrioritizedSampler`` (torchrl/data/replay_buffers/samplers.py:686-1096) on top of a pair of
trees that may be either ``OracleTree`` objects or the compiled reference pybind trees
(``oracle/_ref``: ``SumSegmentTreeFp32`` / ``MinSegmentTreeFp32``) -- the two expose the same
methods (``query``, ``scan_lower_bound``, ``__getitem__``, ``__setitem__``), which is exactly the
surface the reference sampler uses.

All arithmetic that the reference performs with torch CPU ops (``torch.rand * p_sum``,
``torch.pow``) is performed with the same torch CPU ops here, so the oracle's floats are the
reference's floats.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import lib


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class OracleTree:
    """C-oracle twin of ``SumSegmentTreeFp32`` / ``MinSegmentTreeFp32`` (csrc/segment_tree.h:266-307)."""

    def __init__(self, size: int, is_min: bool = False):
        self._L = lib()
        self._t = self._L.orc_tree_new(int(size), int(bool(is_min)))
        self.is_min = bool(is_min)

    def __del__(self):
        try:
            self._L.orc_tree_free(self._t)
        except Exception:
            pass

    # -- properties (segment_tree.h:51-55, pybind :318-322)
    @property
    def size(self) -> int:
        return self._L.orc_tree_size(self._t)

    @property
    def capacity(self) -> int:
        return self._L.orc_tree_capacity(self._t)

    @property
    def identity_element(self) -> float:
        return float(np.finfo(np.float32).max) if self.is_min else 0.0

    def __len__(self) -> int:
        return self.size

    def values(self) -> np.ndarray:
        """The whole implicit heap, ``2*capacity`` fp32 values (index 0 unused)."""
        n = 2 * self.capacity
        buf = (ctypes.c_float * n).from_address(self._L.orc_tree_values(self._t))
        return np.frombuffer(buf, dtype=np.float32, count=n).copy()

    # -- At / __getitem__ (segment_tree.h:56-79)
    def __getitem__(self, index):
        idx = np.ascontiguousarray(np.asarray(index, dtype=np.int64).reshape(-1))
        out = np.empty(idx.shape, dtype=np.float32)
        self._L.orc_tree_at(self._t, _ptr(idx), _ptr(out), idx.size)
        if np.ndim(index) == 0:
            return float(out[0])
        return out.reshape(np.shape(index))

    at = __getitem__

    # -- Update / __setitem__ (segment_tree.h:83-139)
    def __setitem__(self, index, value):
        idx = np.ascontiguousarray(np.asarray(index, dtype=np.int64).reshape(-1))
        val = np.ascontiguousarray(np.asarray(value, dtype=np.float32).reshape(-1))
        if not (val.size == 1 or val.size == idx.size):
            raise ValueError("value must have one element or as many as index")
        self._L.orc_tree_update(self._t, _ptr(idx), _ptr(val), idx.size, int(val.size == 1))

    update = __setitem__

    # -- Query (segment_tree.h:143-162; walk=True -> cuda_segment_tree.cu:51-73, no root fast path)
    def query(self, l, r, walk: bool = False):
        fn = self._L.orc_tree_query_walk if walk else self._L.orc_tree_query
        if np.ndim(l) == 0:
            return float(fn(self._t, int(l), int(r)))
        l = np.asarray(l, dtype=np.int64)
        r = np.asarray(r, dtype=np.int64)
        return np.array([fn(self._t, int(a), int(b)) for a, b in zip(l.reshape(-1), r.reshape(-1))],
                        dtype=np.float32).reshape(l.shape)

    # -- ScanLowerBound (segment_tree.h:249-264,289-294)
    def scan_lower_bound(self, value):
        if self.is_min:
            raise AttributeError("scan_lower_bound is a SumSegmentTree method")
        v = np.ascontiguousarray(np.asarray(value, dtype=np.float32).reshape(-1))
        out = np.empty(v.shape, dtype=np.int64)
        self._L.orc_tree_scan_lower_bound(self._t, _ptr(v), _ptr(out), v.size)
        if np.ndim(value) == 0:
            return int(out[0])
        return out.reshape(np.shape(value))

    def dump_leaves(self) -> np.ndarray:
        out = np.empty(self.size, dtype=np.float32)
        self._L.orc_tree_dump_leaves(self._t, _ptr(out))
        return out

    def load_leaves(self, leaves) -> None:
        leaves = np.ascontiguousarray(np.asarray(leaves, dtype=np.float32))
        assert leaves.size == self.size
        self._L.orc_tree_load_leaves(self._t, _ptr(leaves))


def per_sample_c(sum_tree: OracleTree, min_tree: OracleTree, length: int, u: np.ndarray, beta: float,
                 cpu_checks: bool = True, walk_query: bool = False):
    """All-C PER sample given the uniform draws (orc_per_sample) + the reference's torch.pow.

    Returns (index int64[B], priority_weight fp32[B], p_sum, p_min).
    """
    L = lib()
    u = np.ascontiguousarray(np.asarray(u, dtype=np.float32))
    B = u.size
    index = np.empty(B, dtype=np.int64)
    leaf = np.empty(B, dtype=np.float32)
    p_sum = ctypes.c_float()
    p_min = ctypes.c_float()
    rc = L.orc_per_sample(sum_tree._t, min_tree._t, int(length), _ptr(u), B, int(cpu_checks),
                          int(walk_query), _ptr(index), _ptr(leaf), ctypes.byref(p_sum),
                          ctypes.byref(p_min))
    if rc == -1:
        raise RuntimeError("non-positive p_sum")
    if rc == -2:
        raise RuntimeError("non-positive p_min")
    if rc == -3:
        raise RuntimeError("Failed to find a suitable index")
    # samplers.py:953 -- torch.pow(weight / p_min, -beta), fp32
    w = torch.pow(torch.from_numpy(leaf) / p_min.value, -beta).numpy()
    return index, w, p_sum.value, p_min.value


class OraclePrioritizedSampler:
    """Line-by-line restatement of the reference sampler's arithmetic for 1-D storages.

    ``tree_factory(size, is_min)`` returns a tree object; by default the C oracle tree, but
    ``ref_loader.reference_trees`` can be passed to drive the compiled reference instead.
    Mirrors samplers.py: __init__ :686-714, default_priority :886-893, sample :895-956,
    update_priority :966-1091 (1-D index branch), mark_update :1093-1096.
    """

    def __init__(self, max_capacity: int, alpha: float, beta: float, eps: float = 1e-8,
                 tree_factory=None):
        if alpha < 0:
            raise ValueError(f"alpha must be greater or equal than 0, got alpha={alpha}")
        if beta < 0:
            raise ValueError(f"beta must be greater or equal to 0, got beta={beta}")
        self._max_capacity = int(max_capacity)
        self._alpha, self._beta, self._eps = alpha, beta, eps
        factory = tree_factory or (lambda size, is_min: OracleTree(size, is_min))
        self._sum_tree = factory(self._max_capacity, False)
        self._min_tree = factory(self._max_capacity, True)
        self._max_priority = None  # raw (pre-pow) running max, samplers.py:1054-1075

    @property
    def default_priority(self):
        # samplers.py:886-893: (max_priority + eps) ** alpha, with max_priority = 1 before any update
        mp = self._max_priority
        if mp is None:
            mp = 1
        return (mp + self._eps) ** self._alpha

    def sample(self, length: int, batch_size: int, generator: torch.Generator | None = None,
               u: torch.Tensor | None = None):
        """Returns (index LongTensor[B], priority_weight FloatTensor[B]).

        Either a CPU ``generator`` (mass = torch.rand(B, generator) * p_sum, samplers.py:923) or
        explicit uniform draws ``u`` (same multiply) must be given; parity runs never use the
        np.random branch (:921).
        """
        if length == 0:
            raise RuntimeError("Cannot sample from an empty storage.")
        p_sum = self._sum_tree.query(0, length)
        p_min = self._min_tree.query(0, length)
        if p_sum <= 0:
            raise RuntimeError("non-positive p_sum")
        if p_min <= 0:
            raise RuntimeError("non-positive p_min")
        if u is None:
            u = torch.rand(batch_size, generator=generator)
        mass = u.to(torch.float32) * p_sum  # python float holding an fp32 value -> fp32 multiply
        index = torch.as_tensor(self._sum_tree.scan_lower_bound(mass.numpy()))
        if not index.ndim:
            index = index.unsqueeze(0)
        index = index.clone()
        index.clamp_max_(length - 1)
        weight = torch.as_tensor(self._sum_tree[index.numpy()])
        zero_weight = weight == 0
        while zero_weight.any():
            index = torch.where(zero_weight, index - 1, index)
            if (index < 0).any():
                raise RuntimeError("Failed to find a suitable index")
            weight = torch.as_tensor(self._sum_tree[index.numpy()])
            zero_weight = weight == 0
        weight = torch.pow(weight / p_min, -self._beta)
        return index, weight

    @torch.no_grad()
    def update_priority(self, index, priority) -> None:
        priority = torch.as_tensor(priority).detach()
        index = torch.as_tensor(index, dtype=torch.long)
        if priority.numel() > 1 and priority.shape != index.shape:
            priority = priority.reshape(index.shape[:1])
        elif priority.numel() <= 1:
            priority = priority.squeeze()
        if index.ndim == 0:
            index = index.view(1)
            if priority.ndim == 0:
                priority = priority.view(1)
        valid = index >= 0
        if not valid.any():
            return
        if not valid.all():
            index = index[valid]
            if priority.ndim:
                priority = priority[valid]
        max_p = priority.max(dim=0)[0] if priority.ndim else priority
        if self._max_priority is None or max_p > self._max_priority:
            self._max_priority = max_p
        priority = torch.pow(priority + self._eps, self._alpha)
        pr = priority.to(torch.float32).numpy()
        self._sum_tree[index.numpy()] = pr
        self._min_tree[index.numpy()] = pr

    def mark_update(self, index) -> None:
        self.update_priority(index, self.default_priority)


def gather_rows(src: np.ndarray, index: np.ndarray, length: int) -> np.ndarray:
    """``storage[:len][index]`` for one leaf through the byte-level C restatement."""
    L = lib()
    src = np.ascontiguousarray(src)
    index = np.ascontiguousarray(np.asarray(index, dtype=np.int64))
    row_bytes = src.dtype.itemsize * int(np.prod(src.shape[1:], dtype=np.int64))
    out = np.empty((index.size, *src.shape[1:]), dtype=src.dtype)
    rc = L.orc_gather_rows(_ptr(src), row_bytes, row_bytes, int(length), _ptr(index), index.size, _ptr(out))
    if rc != 0:
        raise IndexError("index out of range")
    return out


def gae_f32(gamma, lmbda, state_value, next_state_value, reward, done, terminated=None):
    """fp32 time-loop GAE (functional.py:119-180) through the C oracle on [*B, T, F] tensors."""
    return _gae(gamma, lmbda, state_value, next_state_value, reward, done, terminated, False)


def gae_f64(gamma, lmbda, state_value, next_state_value, reward, done, terminated=None):
    """float64 evaluation of the same recurrence on the fp32 inputs (accuracy ground truth)."""
    return _gae(gamma, lmbda, state_value, next_state_value, reward, done, terminated, True)


def _gae(gamma, lmbda, v, nv, r, done, term, f64: bool):
    L = lib()
    if term is None:
        term = done
    shape = tuple(v.shape)
    assert len(shape) >= 2, "expected [*B, T, F]"
    T, F = shape[-2], shape[-1]
    rows = int(np.prod(shape[:-2], dtype=np.int64)) if len(shape) > 2 else 1
    g32 = torch.as_tensor(gamma, dtype=torch.float32)
    l32 = torch.as_tensor(lmbda, dtype=torch.float32)
    c = lambda t, dt: np.ascontiguousarray(t.detach().cpu().numpy().astype(dt, copy=False))
    v_, nv_, r_ = c(v, np.float32), c(nv, np.float32), c(r, np.float32)
    d_, t_ = c(done, np.uint8), c(term, np.uint8)
    if f64:
        adv = np.empty(shape, dtype=np.float64)
        tgt = np.empty(shape, dtype=np.float64)
        # the fp32-rounded gamma and gamma*lmbda the reference actually uses, evaluated in fp64
        gl = float(g32 * l32)
        L.orc_gae_f64(_ptr(v_), _ptr(nv_), _ptr(r_), _ptr(d_), _ptr(t_), float(g32), gl, rows, T, F,
                      _ptr(adv), _ptr(tgt))
    else:
        adv = np.empty(shape, dtype=np.float32)
        tgt = np.empty(shape, dtype=np.float32)
        gl = float(l32 * g32)  # functional.py:172 `lmbda * gamma` as fp32 0-d tensors
        L.orc_gae_f32(_ptr(v_), _ptr(nv_), _ptr(r_), _ptr(d_), _ptr(t_), float(g32), gl, rows, T, F,
                      _ptr(adv), _ptr(tgt))
    return torch.from_numpy(adv), torch.from_numpy(tgt)


def td_lambda(gamma, lmbda, next_state_value, reward, done, terminated=None, f64: bool = False):
    """TD(lambda) return (functional.py:790-899, scalar gamma/lmbda) through the C oracle; fp32 (reference op order)
    or float64 evaluation of the same recurrence on the fp32 inputs with the fp32-rounded gamma / lmbda."""
    L = lib()
    if terminated is None:
        terminated = done
    shape = tuple(reward.shape)
    T, F = shape[-2], shape[-1]
    rows = int(np.prod(shape[:-2], dtype=np.int64)) if len(shape) > 2 else 1
    g32 = float(torch.as_tensor(gamma, dtype=torch.float32))
    l32 = float(torch.as_tensor(lmbda, dtype=torch.float32))
    c = lambda t, dt: np.ascontiguousarray(t.detach().cpu().numpy().astype(dt, copy=False))
    nv_, r_, d_, t_ = c(next_state_value, np.float32), c(reward, np.float32), c(done, np.uint8), c(terminated, np.uint8)
    if f64:
        out = np.empty(shape, dtype=np.float64)
        L.orc_td_lambda_f64(_ptr(nv_), _ptr(r_), _ptr(d_), _ptr(t_), g32, l32, rows, T, F, _ptr(out))
    else:
        out = np.empty(shape, dtype=np.float32)
        L.orc_td_lambda_f32(_ptr(nv_), _ptr(r_), _ptr(d_), _ptr(t_), g32, l32, rows, T, F, _ptr(out))
    return torch.from_numpy(out)


def affine_scan(d: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """out_t = d_t + c_t * out_{t+1} along dim -2 of [*B, T, F] tensors, through the C oracle (dtype of ``d``)."""
    L = lib()
    shape = tuple(d.shape)
    T, F = shape[-2], shape[-1]
    rows = int(np.prod(shape[:-2], dtype=np.int64)) if len(shape) > 2 else 1
    dt = np.float64 if d.dtype == torch.float64 else np.float32
    d_ = np.ascontiguousarray(d.detach().cpu().numpy().astype(dt, copy=False))
    c_ = np.ascontiguousarray(torch.broadcast_to(c, d.shape).detach().cpu().numpy().astype(dt, copy=False))
    out = np.empty(shape, dtype=dt)
    (L.orc_affine_scan_f64 if dt == np.float64 else L.orc_affine_scan_f32)(_ptr(d_), _ptr(c_), rows, T, F, _ptr(out))
    return torch.from_numpy(out)


def vtrace(gamma, log_pi, log_mu, state_value, next_state_value, reward, done, terminated=None, rho_thresh=1.0,
           c_thresh=1.0):
    """V-trace (vtrace_advantage_estimate, functional.py:1297-1382), time at dim -2: the reference's elementwise
    prologue / epilogue as CPU tensor ops in its order, its python time loop (:1360-1368) through the C scan."""
    rho_thresh = torch.as_tensor(rho_thresh)
    c_thresh = torch.as_tensor(c_thresh)
    not_done = (~done).int()
    not_terminated = not_done if terminated is None else (~terminated).int()
    done_discounts = gamma * not_done
    terminated_discounts = gamma * not_terminated
    rho = (log_pi - log_mu).exp()
    clipped_rho = rho.clamp_max(rho_thresh)
    deltas = clipped_rho * (reward + terminated_discounts * next_state_value - state_value)
    clipped_c = rho.clamp_max(c_thresh)
    vs_minus_v = affine_scan(deltas, done_discounts * clipped_c)
    vs = vs_minus_v + state_value
    vs_t_plus_1 = torch.cat([vs[..., 1:, :], next_state_value[..., -1:, :]], dim=-2)
    advantages = clipped_rho * (reward + terminated_discounts * vs_t_plus_1 - state_value)
    return advantages, vs


def gae_per_step(gamma, lmbda, state_value, next_state_value, reward, done, terminated=None):
    """GAE with tensor-valued gamma / lmbda (functional.py:317-370), time at dim -2, as the recurrence the rolled
    cumprod tensor (value/utils.py:130-181) unrolls to:  A_t = td0_t + not_done_t*gamma_t*lmbda_t * A_{t+1}."""
    if terminated is None:
        terminated = done
    dtype = state_value.dtype
    value = gamma * lmbda
    gammalmbdas = (~done).to(dtype) * value
    td0 = reward + (~terminated).to(dtype) * gamma * next_state_value - state_value
    adv = affine_scan(td0, gammalmbdas)
    return adv, adv + state_value
