"""TESTS ONLY -- an oracle-backed stand-in for ``rl_b200.ops.CudaBackend``.

Lets the CPU test-suite drive the *host logic* of rl_b200 (cursors, lengths, key plumbing, bookkeeping,
error behaviour, multi-process sharding) on CPU tensors.  Every method does what the corresponding C-ABI
entry point is specified to do (include/rlb200.h), computed with the CPU oracle (oracle/) and plain torch
indexing.  It is never installed by product code and never used on the GPU box: the ``-m gpu`` tests call
the real library.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import lib as orc_lib
from oracle import framestack_oracle as fo
from oracle import per_oracle as po


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().numpy()


class OracleBackend:
    name = "oracle-emulator"

    # ---- segment tree: the heap tensor is kept bit-identical to what the kernels must produce
    def tree_capacity(self, size: int) -> int:
        return _capacity(size)

    def l2_persist(self, tensor, stream=None):
        return 0

    def tree_new(self, size, is_min, dtype, device, out=None):
        cap = _capacity(size)
        ident = torch.finfo(dtype).max if is_min else 0.0
        if out is not None:
            return out.fill_(ident)
        return torch.full((2 * cap,), ident, dtype=dtype, device="cpu")

    def tree_workspace(self, size, device):
        return torch.zeros(_capacity(size), dtype=torch.int64)

    def tree_rebuild(self, tree, capacity, is_min):
        for i in range(capacity - 1, 0, -1):
            a, b = tree[2 * i], tree[2 * i + 1]
            tree[i] = torch.minimum(a, b) if is_min else a + b

    def _update_one(self, tree, capacity, is_min, index, value):
        if tree is None:
            return
        idx = index.tolist()
        vals = value.tolist() if value.numel() > 1 else [value.item()] * len(idx)
        for i, v in zip(idx, vals):
            if i < 0:
                continue
            node = i + capacity
            tree[node] = v
            while node > 1:
                p = node >> 1
                a, b = tree[2 * p], tree[2 * p + 1]
                tree[p] = torch.minimum(a, b) if is_min else a + b
                node = p

    def tree_update(self, sum_tree, min_tree, capacity, index, value, workspace, epoch):
        ref = sum_tree if sum_tree is not None else min_tree
        if value.dtype != ref.dtype:
            raise RuntimeError("value dtype must match the tree dtype")
        if index.dtype != torch.int64:
            raise RuntimeError("index must be an int64 tensor")
        self._update_one(sum_tree, capacity, False, index.reshape(-1), value.reshape(-1))
        self._update_one(min_tree, capacity, True, index.reshape(-1), value.reshape(-1))

    def _query1(self, tree, size, capacity, is_min, l, r, root_fast_path):
        if root_fast_path and l <= 0 and r >= size:
            return tree[1].clone()
        ret = torch.tensor(torch.finfo(tree.dtype).max if is_min else 0.0, dtype=tree.dtype)
        l |= capacity
        r |= capacity
        while l < r:
            if l & 1:
                ret = torch.minimum(ret, tree[l]) if is_min else ret + tree[l]
                l += 1
            if r & 1:
                r -= 1
                ret = torch.minimum(ret, tree[r]) if is_min else ret + tree[r]
            l >>= 1
            r >>= 1
        return ret

    def tree_query(self, tree, size, capacity, is_min, l, r, root_fast_path):
        out = [self._query1(tree, size, capacity, is_min, int(a), int(b), root_fast_path)
               for a, b in zip(l.reshape(-1).tolist(), r.reshape(-1).tolist())]
        return torch.stack(out).reshape(l.shape) if out else torch.empty(l.shape, dtype=tree.dtype)

    def tree_at(self, tree, capacity, index):
        return tree[index + capacity]

    def _scan1(self, tree, size, capacity, value):
        if value > tree[1]:
            return size
        node, cur = 1, value.clone()
        while node < capacity:
            node <<= 1
            if cur > tree[node]:
                cur = cur - tree[node]
                node |= 1
        return node ^ capacity

    def tree_scan_lower_bound(self, tree, size, capacity, value):
        if value.dtype != tree.dtype:
            raise RuntimeError("value dtype must match the tree dtype")
        out = [self._scan1(tree, size, capacity, v) for v in value.reshape(-1)]
        return torch.tensor(out, dtype=torch.int64).reshape(value.shape)

    # ---- fused sampler arithmetic (restates include/rlb200.h rlb_per_sample / rlb_per_update)
    def per_sample(self, sum_tree, min_tree, size, capacity, length, u, beta, cpu_semantics, status=None,
                   want_aux=False, out=None):
        p_sum = self._query1(sum_tree, size, capacity, False, 0, length, bool(cpu_semantics))
        p_min = self._query1(min_tree, size, capacity, True, 0, length, bool(cpu_semantics))
        if status is not None:
            if not p_sum > 0:
                status |= 2
            if not p_min > 0:
                status |= 4
        mass = u * p_sum
        index = self.tree_scan_lower_bound(sum_tree, size, capacity, mass)
        index = index.clamp_max(length - 1)
        leaf = sum_tree[index + capacity]
        if cpu_semantics:
            zero = leaf == 0
            while zero.any():
                index = torch.where(zero, index - 1, index)
                if (index < 0).any():
                    if status is not None:
                        status |= 8
                    index = index.clamp_min(0)
                    break
                leaf = sum_tree[index + capacity]
                zero = leaf == 0
        weight = torch.pow(leaf / p_min, -beta).to(torch.float32)
        if out is not None:
            for dst, src in zip(out, (index, weight, leaf, torch.stack([p_sum, p_min]))):
                dst.copy_(src)
            return out
        if want_aux:
            return index, weight, leaf, torch.stack([p_sum, p_min])
        return index, weight

    def per_update(self, sum_tree, min_tree, capacity, index, priority, alpha, eps, max_out, workspace, epoch,
                   index_base=0, index_limit=-1):
        index = index.reshape(-1) - index_base
        limit = capacity if index_limit < 0 else index_limit
        index = torch.where((index >= 0) & (index < limit), index, torch.full_like(index, -1))
        priority = priority.reshape(-1).to(torch.float32)
        valid = index >= 0
        if max_out is not None and valid.any():
            pv = priority.expand_as(index)[valid] if priority.numel() == 1 else priority[valid]
            max_out.copy_(torch.maximum(max_out, pv.max().view(1)))
        leaf = torch.pow(priority + eps, alpha)
        self.tree_update(sum_tree, min_tree, capacity, index, leaf, workspace, epoch)

    # ---- write path: what the reference does for a writer batch, index by index
    def tree_update_range(self, rng, start, n, modulo):
        index = torch.arange(start, start + n) % modulo
        if rng.mode == 0:
            return self.tree_update(rng.sum, rng.mn, rng.capacity, index, rng.value.reshape(1), None, 0)
        if rng.mode == 1:
            p = rng.value.reshape(()).to(torch.float32)
        elif rng.has_max and bool(rng.max_buf[0] > float("-inf")):
            p = (rng.max_buf[0] + rng.eps) ** rng.alpha                     # samplers.py:886-893
        else:
            p = torch.as_tensor(rng.first_default, dtype=torch.float32)
        self.per_update(rng.sum, rng.mn, rng.capacity, index, p, rng.alpha, rng.eps, rng.max_buf, None, 0)

    def extend(self, stores, data, cursor, n, max_size, rng=None):
        index = torch.arange(cursor, cursor + n) % max_size
        for t, d in zip(stores, data):
            if t.numel() and d.numel():
                t[index] = d
        if rng is not None:
            self.tree_update_range(rng, cursor, n, max_size)

    # ---- trajectory slices: the restated reference arithmetic (oracle/slice_oracle.py)
    def traj_workspace(self, L, device):
        return torch.zeros(1, dtype=torch.int64)

    def traj_table(self, signal, by_id, L, at_capacity, cursor, min_len, keep_long_only, table, counts, workspace):
        from oracle import slice_oracle as so

        sig = signal.reshape(-1)[:L].numpy()
        cur = None if cursor < 0 else cursor
        start, stop, length = (so.traj_table(trajectory=sig, at_capacity=at_capacity, cursor=cur) if by_id
                               else so.traj_table(end=sig, at_capacity=at_capacity, cursor=cur))
        long_enough = length >= min_len
        counts[0], counts[1] = len(start), int(long_enough.sum())
        if keep_long_only:
            start, stop, length = start[long_enough], stop[long_enough], length[long_enough]
        for row, a in zip(table, (start, stop, length)):
            row[:len(a)] = torch.from_numpy(a)

    def slice_index(self, start, length, n_traj, traj_draw, u, seq_length, storage_length, variable=False,
                    pad_output=False, out_offset=None, total=None, want_index=True, flags=None, span=(0, 0)):
        from oracle import slice_oracle as so

        idx, tr, mask, seq = so.slice_index(start[:n_traj].numpy(), length[:n_traj].numpy(), seq_length=seq_length,
                                            num_slices=traj_draw.numel(), storage_length=storage_length,
                                            traj_draw=traj_draw.numpy(), u=u.numpy(), strict_length=not variable,
                                            pad_output=pad_output, span=span, force_variable=variable)
        seq = torch.from_numpy(np.asarray(seq))
        if not want_index:
            return None, None, None, seq
        out = (torch.from_numpy(idx), torch.from_numpy(tr).reshape(-1, 1),
               None if mask is None else torch.from_numpy(mask), seq)
        if flags is None:
            return out
        index, trunc = out[0], out[1]
        done = trunc.clone() if flags[0] is None else flags[0].reshape(-1)[index].reshape(-1, 1).bool() | trunc
        term = torch.zeros_like(trunc) if flags[1] is None else flags[1].reshape(-1)[index].reshape(-1, 1).bool()
        return (*out, done, term)

    def slice_mask_starts(self, masked_tree, capacity, stop, length, n_traj, seq_length, ring_length):
        from oracle import slice_oracle as so

        bad = so.invalid_starts(stop[:n_traj].numpy(), length[:n_traj].numpy(), seq_length, ring_length)
        masked_tree[capacity + torch.from_numpy(bad)] = 0

    # ---- sharded minibatch trailer
    def shard_pack(self, rows, meta_offset, index, leaf, psum_pmin, index_base, peer_delta=None, flags=None,
                   seq_counter=None, rank=0):
        m = meta_offset
        rows[:, m:m + 8].view(torch.int64).view(-1).copy_(index + index_base)
        rows[:, m + 8:m + 12].view(torch.float32).view(-1).copy_(leaf)
        rows[:, m + 12:m + 16].view(torch.float32).view(-1).copy_(psum_pmin[0].expand(rows.shape[0]))
        rows[:, m + 16:m + 20].view(torch.float32).view(-1).copy_(psum_pmin[1].expand(rows.shape[0]))

    def shard_weights(self, rows, meta_offset, beta, flags=None, wait_counter=None, n_ranks=0, timeout_s=10.0,
                      status=None, out=None):
        m = meta_offset
        gidx = rows[:, m:m + 8].view(torch.int64).view(-1).clone()
        p = rows[:, m + 8:m + 12].view(torch.float32).view(-1)
        S = rows[:, m + 12:m + 16].view(torch.float32).view(-1)
        mn = rows[:, m + 16:m + 20].view(torch.float32).view(-1)
        return torch.pow((p / S) / (mn / S).min(), -beta), gidx

    # ---- storage rows
    def gather(self, leaves, index, length, mode=0, status=None, out=None, peer_delta=None, multicast_delta=0):
        ix = torch.where(index < 0, index + length, index)
        if ((ix < 0) | (ix >= length)).any():
            if status is not None:
                status |= 1
            ix = ix.clamp(0, length - 1)
        res = [t[ix] for t in leaves]
        if out is not None:
            for o, r in zip(out, res):
                o.copy_(r)
            return list(out)
        return res

    def gather_plan(self, leaves, frames=None):
        be = self

        class _Plan:
            def run(self, index, length, mode=0, status=None, out=None, peer_delta=None, multicast_delta=0):
                if frames is None or not any(f is not None for f in frames):
                    return be.gather(leaves, index, length, mode=mode, status=status, out=out)
                res = []
                for k, (leaf, f) in enumerate(zip(leaves, frames)):
                    if f is None:
                        r = be.gather([leaf], index, length, status=status)[0]
                    else:   # rlb_gather_ex: slot -> frame word -> pool row
                        word, head, ring, off = f
                        w = be.gather([word], index, length, status=status)[0]
                        env, q = w >> fo.ENV_SHIFT, (w & fo.POS_MASK) + off
                        if head is not None and status is not None and ((q < 0) | (head[env] - q > ring)).any():
                            status |= 32
                        r = leaf[env * ring + q.clamp(min=0) % ring]
                    if out is not None:
                        out[k].copy_(r)
                    res.append(r)
                return list(out) if out is not None else res

        return _Plan()

    def framestack_push(self, obs, next_obs, is_init, done, last_done, head, pool, n_envs, layout, k, ring):
        pl, hd, ld = pool.numpy(), head.numpy(), last_done.numpy()   # CPU tensors: numpy views, mutated in place
        words = fo.push(pl, hd, ld, obs.numpy(), next_obs.numpy(), None if is_init is None else is_init.numpy(),
                        None if done is None else done.numpy(), n_envs=n_envs, layout=layout, k=k, ring=ring)
        return torch.from_numpy(words)

    def scatter(self, leaves, data, index, length, status=None):
        ix = torch.where(index < 0, index + length, index)
        for t, d in zip(leaves, data):
            t[ix] = d

    # ---- GAE
    def gae(self, v, nv, r, done, term, gamma, gammalmbda, rows, T, F):
        L = orc_lib()
        shape = v.shape
        if v.dtype == torch.float32:
            adv = np.empty((rows, T, F), dtype=np.float32)
            tgt = np.empty_like(adv)
            a = [np.ascontiguousarray(_np(x)) for x in (v, nv, r, done, term)]
            L.orc_gae_f32(*[x.ctypes.data for x in a], gamma, gammalmbda, rows, T, F, adv.ctypes.data,
                          tgt.ctypes.data)
        else:
            adv = np.empty((rows, T, F), dtype=np.float64)
            tgt = np.empty_like(adv)
            a = [np.ascontiguousarray(_np(x).astype(np.float32)) for x in (v, nv, r)] + \
                [np.ascontiguousarray(_np(x)) for x in (done, term)]
            L.orc_gae_f64(*[x.ctypes.data for x in a], gamma, gammalmbda, rows, T, F, adv.ctypes.data,
                          tgt.ctypes.data)
        return torch.from_numpy(adv).reshape(shape), torch.from_numpy(tgt).reshape(shape)


    def td_lambda_return(self, nv, r, done, term, gamma, gammalmbda, one_minus_lmbda, rows, T, F):
        lmbda = 1.0 - one_minus_lmbda
        return po.td_lambda(gamma, lmbda, nv, r, done.bool(), term.bool(), f64=(nv.dtype == torch.float64)).to(nv.dtype)


    def affine_scan(self, d, c, rows, T, F):
        return po.affine_scan(d.reshape(rows, T, F), c.reshape(rows, T, F)).reshape(d.shape)


def _capacity(size: int) -> int:
    c = 1
    while c <= size:
        c <<= 1
    return c
