"""TEST INFRASTRUCTURE -- CPU restatement of the de-duplicated frame-stack storage (include/rlb200.h: rlb_framestack_push,
rlb_gather_ex).

The reference has no such storage: what it does with the same transitions is keep both stacks of every transition verbatim
(TensorStorage.set / get, torchrl/data/replay_buffers/storages.py:1028-1096 and :1098-1130).  The parity statement of the
de-duplicated storage is therefore "a batch read back equals the batch the reference's TensorStorage returns for the same
indices", and this file holds (a) ``MaterialisedStorage`` -- that reference behaviour (rows in, the same rows out) in numpy;
the reference class itself cannot be imported here (it needs the absent ``tensordict`` package), so this half is a
restatement of ``storage[index] = data`` / ``storage[index]``, nothing more -- and (b) a plain-loop restatement of the frame
log itself, used to check the kernels' frame words, ring heads and pool bytes.  Never imported by the product.
"""
from __future__ import annotations

import numpy as np

ENV_SHIFT = 40
POS_MASK = (1 << ENV_SHIFT) - 1


class MaterialisedStorage:
    """What the reference stores: both stacks of every transition, rows addressed by slot (storages.py:1028-1130)."""

    def __init__(self, max_size: int):
        self.max_size, self.obs, self.next, self.len = max_size, None, None, 0

    def set(self, slots, obs, nxt):
        if self.obs is None:
            self.obs = np.zeros((self.max_size, *obs.shape[1:]), dtype=obs.dtype)
            self.next = np.zeros_like(self.obs)
        self.obs[slots] = obs
        self.next[slots] = nxt
        self.len = min(self.len + len(slots), self.max_size)

    def get(self, index):
        return self.obs[index], self.next[index]


def row_of(env: int, step: int, steps: int, n_envs: int, layout: int) -> int:
    return env * steps + step if layout == 0 else step * n_envs + env


def push(pool, head, last_done, obs, nxt, is_init, done, *, n_envs: int, layout: int, k: int, ring: int):
    """One rlb_framestack_push call, environment by environment, step by step.  Mutates pool / head / last_done; returns
    the frame words int64[n]."""
    n = obs.shape[0]
    steps = n // n_envs
    words = np.zeros(n, dtype=np.int64)
    for env in range(n_envs):
        pos = int(head[env])
        for t in range(steps):
            i = row_of(env, t, steps, n_envs, layout)
            if is_init is not None:
                start = bool(is_init[i])
            else:
                start = bool(last_done[env]) if t == 0 else bool(done[row_of(env, t - 1, steps, n_envs, layout)])
            if start:
                for j in range(k):
                    pool[env * ring + pos % ring] = obs[i, j]
                    pos += 1
            pool[env * ring + pos % ring] = nxt[i, k - 1]
            words[i] = (env << ENV_SHIFT) | pos
            pos += 1
        head[env] = pos
        if done is not None:
            last_done[env] = bool(done[row_of(env, steps - 1, steps, n_envs, layout)])
        elif is_init is not None:
            last_done[env] = 0
    return words


def rebuild(pool, head, words, *, k: int, ring: int):
    """(obs [B, k, ...], next [B, k, ...], evicted bool[B]) for the transitions with frame words ``words``."""
    B = len(words)
    obs = np.zeros((B, k, *pool.shape[1:]), dtype=pool.dtype)
    nxt = np.zeros_like(obs)
    evicted = np.zeros(B, dtype=bool)
    for b, w in enumerate(words):
        env, p = int(w) >> ENV_SHIFT, int(w) & POS_MASK
        evicted[b] = int(head[env]) - (p - k) > ring
        for j in range(k):
            obs[b, j] = pool[env * ring + (p - k + j) % ring]
            nxt[b, j] = pool[env * ring + (p - k + 1 + j) % ring]
    return obs, nxt, evicted


def make_stream(n_envs: int, steps: int, k: int, frame=(6, 5), *, seed: int = 0, pad: str = "same", min_len: int = 1,
                max_len: int = 40, dtype=np.uint8):
    """A synthetic frame-stacked stream the way a CatFrames-wrapped vectorised environment produces it:
    (obs [steps, E, k, *frame], next [...], done bool[steps, E], is_init bool[steps, E]).
    ``pad``: how the stack is filled at a reset -- "same" repeats the first frame, "constant" pads with zeros."""
    rng = np.random.default_rng(seed)
    obs = np.zeros((steps, n_envs, k, *frame), dtype=dtype)
    nxt = np.zeros_like(obs)
    done = np.zeros((steps, n_envs), dtype=bool)
    init = np.zeros((steps, n_envs), dtype=bool)

    def new_frame():
        return rng.integers(0, 255, size=frame).astype(dtype)

    for e in range(n_envs):
        stack, left = None, 0
        for t in range(steps):
            if stack is None:
                f0 = new_frame()
                stack = np.stack([f0] * k) if pad == "same" else np.stack([np.zeros(frame, dtype)] * (k - 1) + [f0])
                left = int(rng.integers(min_len, max_len + 1))
                init[t, e] = True
            obs[t, e] = stack
            stack = np.concatenate([stack[1:], new_frame()[None]])
            nxt[t, e] = stack
            left -= 1
            if left == 0:
                done[t, e] = True
                stack = None
    return obs, nxt, done, init
