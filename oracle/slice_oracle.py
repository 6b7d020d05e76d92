"""TEST INFRASTRUCTURE -- CPU restatement of the reference's SliceSampler arithmetic (1-d storages; N-d ring by ring).

Follows torchrl/data/replay_buffers/samplers.py (SliceSampler):
    traj_table    _find_start_stop_traj :1652-1706  +  _end_to_start_stop :1708-1743
    slice_index   _sample_slices :1973-2056  +  _get_index :2058-2215  (incl. span)
with the random draws passed in explicitly (``traj_draw`` = the output of ``torch.randint(maxval, (num_slices,))``,
``u`` = the output of ``torch.rand(num_slices)``; the reference makes exactly these two calls in this order, :1987-1990 and
:2099-2102), so that a CUDA run can be checked against it with the draws its own generator produced.
Pinned against the unmodified reference (tests/test_oracle.py: patched RNG calls) and tests/golden/slice_golden.npz.
Never imported by the product.
"""
from __future__ import annotations

import numpy as np


def traj_table(*, end=None, trajectory=None, at_capacity: bool, cursor=None):
    """(start, stop, length) int64 arrays, one entry per trajectory in the ring, ordered by stop position."""
    if trajectory is not None:
        t = np.asarray(trajectory)
        L = t.shape[0]
        end = np.empty(L, dtype=bool)
        end[:-1] = t[:-1] != t[1:]                                   # :1666
        end[-1] = (t[-1] != t[0]) if at_capacity else True           # :1667-1670
    else:
        end = np.array(end, dtype=bool).reshape(-1).copy()
        L = end.shape[0]
        if not at_capacity:
            end[L - 1] = True                                        # :1675-1677
    if at_capacity:
        if cursor is not None:
            end[int(cursor)] = True                                  # :1683-1699: the last written slot closes a trajectory
        if not end.any():
            end[L - 1] = True                                        # :1700-1703
    stop = np.nonzero(end)[0].astype(np.int64)                       # :1717
    start = (np.roll(stop, 1) + 1) % L                               # :1720-1738 (1-d: the start is the previous stop + 1)
    length = stop - start + 1                                        # :1739
    length[length <= 0] += L                                         # :1740
    return start, stop, length


def valid_trajectories(start, stop, length, seq_length: int, strict_length: bool):
    """The trajectories `traj_idx` indexes (:1993-2010): with strict_length those at least seq_length long."""
    if strict_length and (length < seq_length).any():
        keep = length >= seq_length
        if not keep.any():
            raise RuntimeError("Did not find a single trajectory with sufficient length")
        return start[keep], stop[keep], length[keep]
    return start, stop, length


def slice_index(start, length, *, seq_length: int, num_slices: int, storage_length: int, traj_draw, u,
                strict_length: bool = True, pad_output: bool = False, span=(0, 0), force_variable: bool = False):
    """index int64[n_out], truncated bool[n_out], mask bool[n_out] | None, per-slice lengths int64[num_slices].

    ``start`` / ``length`` are the (already filtered) trajectories; ``traj_draw`` in [0, len(start)).
    ``span`` = (left, right) as the kernel takes it: 0 off, -1 True, k > 0 (:2071-2118)."""
    traj = np.asarray(traj_draw, dtype=np.int64)
    u = np.asarray(u, dtype=np.float32)
    lens = np.asarray(length, dtype=np.int64)[traj]
    if (not strict_length) and (np.asarray(length) < seq_length).any():
        seq = np.minimum(lens, seq_length)                           # :2037
        variable = True
    else:
        seq = np.full(num_slices, seq_length, dtype=np.int64)
        variable = bool(force_variable)
    span0, span1 = (int(x) for x in span)
    last_indexable_start = lens - seq + 1                            # :2072
    end_point = last_indexable_start if span1 == 0 else (lens + 1 if span1 < 0 else lens - span1)      # :2073-2083
    start_point = np.zeros_like(seq) if span0 == 0 else (1 - seq if span0 < 0 else np.full_like(seq, -span0))  # :2085-2097
    # torch.rand(fp32) * int64 tensor -> fp32 product, floor, cast (:2099-2102)
    rel = np.floor(u * (end_point - start_point).astype(np.float32)).astype(np.int64) + start_point
    if span0:                                                        # :2104-2111
        out = rel < 0
        if out.any():
            seq = np.where(out, seq + rel, seq)
            rel = np.where(out, 0, rel)
            variable = True
    if span1:                                                        # :2112-2118
        out = rel + seq > lens
        if out.any():
            seq = np.minimum(seq, lens - rel)
            variable = True
    seq = np.maximum(seq, 0)                                         # (a trajectory shorter than the span)
    starts = np.asarray(start, dtype=np.int64)[traj] + rel           # :2120-2126
    if variable and pad_output:
        T = seq_length
        ar = np.arange(T, dtype=np.int64)
        real = ar[None, :] < seq[:, None]                            # :2147-2148
        full = starts[:, None] + ar[None, :]
        last = starts + np.maximum(seq - 1, 0)
        full = np.where(real, full, last[:, None]) % storage_length  # :2151-2158
        index = full.reshape(-1)
        mask = real.reshape(-1)
        truncated = np.zeros(num_slices * T, dtype=bool)
        truncated[np.arange(num_slices) * T + np.maximum(seq - 1, 0)] = True      # :2178-2183
        return index, truncated, mask, seq
    if variable:
        index = np.concatenate([s + np.arange(n, dtype=np.int64) for s, n in zip(starts, seq)]) % storage_length
        truncated = np.zeros(index.shape[0], dtype=bool)
        truncated[np.cumsum(seq) - 1] = True                         # :2187
        return index, truncated, None, seq
    index = ((starts[:, None] + np.arange(seq_length, dtype=np.int64)[None, :]) % storage_length).reshape(-1)
    truncated = np.zeros(num_slices * seq_length, dtype=bool)
    truncated.reshape(num_slices, seq_length)[:, -1] = True          # :2185
    return index, truncated, None, seq


def traj_table_nd(*, end=None, trajectory=None, at_capacity: bool, cursor=None):
    """N-d storages ([T, E, ...] signals, time along dim 0): the reference transposes and takes nonzero (:1717-1718), i.e.
    trajectories are listed ring by ring (column by column), each ring exactly as a 1-d storage.  Returns
    (start_t, stop_t, length, column) with the column index of every trajectory."""
    sig = np.asarray(end if end is not None else trajectory)
    T = sig.shape[0]
    cols = sig.reshape(T, -1)
    out = [[], [], [], []]
    for e in range(cols.shape[1]):
        if end is not None:
            st, sp, ln = traj_table(end=cols[:, e], at_capacity=at_capacity, cursor=cursor)
        else:
            st, sp, ln = traj_table(trajectory=cols[:, e], at_capacity=at_capacity, cursor=cursor)
        for o, a in zip(out, (st, sp, ln, np.full(len(st), e, dtype=np.int64))):
            o.append(a)
    return tuple(np.concatenate(o) for o in out)


def invalid_starts(stop, length, seq_length: int, ring_length: int, strict_length: bool = True):
    """PrioritizedSliceSampler._preceding_stop_idx (:2854-2888, span=False): the slots a slice of ``seq_length`` steps must
    not start at -- the last ``seq_length - 1`` steps of every trajectory (all of a shorter one); with
    ``strict_length=False`` the first step of a trajectory is never among them (:2863-2871 removes the starts from the
    candidates first).  The reference lists them through a left-padded index matrix in "trajectory order" coordinates and
    shifts by the first start when the ring is full; in storage coordinates that is simply ``stop - j`` (mod ring) for
    j < min(len, seq - 1) (len - 1 when not strict).
    Order: trajectory by trajectory, ascending within each, like the reference's boolean-mask read-out."""
    out = []
    for sp, ln in zip(np.asarray(stop, dtype=np.int64), np.asarray(length, dtype=np.int64)):
        m = int(min(ln if strict_length else ln - 1, seq_length - 1))
        out.append((sp - np.arange(m - 1, -1, -1, dtype=np.int64)) % ring_length)
    return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)


def prioritized_slice_sample(orc_sampler, start, stop, length, *, seq_length: int, num_slices: int, storage_len: int,
                             u, strict_length: bool = True):
    """PrioritizedSliceSampler.sample (:2890-3004) on an oracle.per_oracle.OraclePrioritizedSampler: zero the invalid
    starts in the sum tree (:2910-2912), draw ``num_slices`` starts like PrioritizedSampler.sample (:2915-2917), restore
    (:2918), cut every slice at the end of its trajectory when not strict (:2919-2951), expand (:2963-2966) and repeat
    the weights (:2969-2971).
    Returns (index int64[n], weight fp32[n], truncated bool[n], starts int64[S])."""
    bad = invalid_starts(stop, length, seq_length, storage_len, strict_length)
    tree = orc_sampler._sum_tree
    vals = np.array(tree[bad], dtype=np.float32)
    tree[bad] = np.zeros(len(bad), dtype=np.float32)
    try:
        import torch

        starts, w = orc_sampler.sample(storage_len, num_slices, u=torch.as_tensor(np.asarray(u)))
    finally:
        tree[bad] = vals
    starts = starts.numpy().astype(np.int64)
    if strict_length:
        index = ((starts[:, None] + np.arange(seq_length, dtype=np.int64)[None, :]) % storage_len).reshape(-1)
        weight = np.repeat(w.numpy(), seq_length)
        truncated = np.zeros(num_slices * seq_length, dtype=bool)
        truncated.reshape(num_slices, seq_length)[:, -1] = True
        return index, weight, truncated, starts
    stop = np.asarray(stop, dtype=np.int64)
    start = np.asarray(start, dtype=np.int64)
    stop_corr = np.where(stop < start, stop + storage_len, stop)                  # :2929-2934
    diff = stop_corr[:, None] - starts[None, :]                                  # :2935
    diff[diff < 0] = diff.max() + 1                                              # :2942
    stops = stop_corr[diff.argmin(axis=0)]                                       # :2944-2945
    seq = np.minimum(stops - starts + 1, seq_length)                             # :2950
    index = np.concatenate([s + np.arange(n, dtype=np.int64) for s, n in zip(starts, seq)]) % storage_len
    weight = np.repeat(w.numpy(), seq)
    truncated = np.zeros(index.shape[0], dtype=bool)
    truncated[np.cumsum(seq) - 1] = True
    return index, weight, truncated, starts
