"""De-duplicated frame-stack storage: a batch read back must equal, bit for bit, what the
reference's TensorStorage returns for the same transitions -- i.e. the stacks that were written (oracle/framestack_oracle.py:
MaterialisedStorage) -- and the frame log (words, heads, pool bytes) must equal the plain-loop restatement."""
import numpy as np
import pytest
import torch

from oracle import framestack_oracle as fo
from rl_b200 import ops
from rl_b200.data import (FrameStackStorage, LazyTensorStorage, RandomSampler, TensorDict, TensorDictPrioritizedReplayBuffer,
                          TensorDictReplayBuffer)

K = 4


def _batches(n_envs, total_steps, steps_per_batch, layout, *, seed, pad, min_len=1, max_len=40, frame=(6, 5), dev="cpu"):
    """Yields (TensorDict batch, obs, next) in writer order from one synthetic stream."""
    obs, nxt, done, init = fo.make_stream(n_envs, total_steps, K, frame, seed=seed, pad=pad, min_len=min_len,
                                          max_len=max_len)
    rng = np.random.default_rng(seed + 1)
    act = rng.integers(0, 18, size=(total_steps, n_envs, 1))
    rew = rng.standard_normal((total_steps, n_envs, 1)).astype(np.float32)
    for t0 in range(0, total_steps, steps_per_batch):
        sl = slice(t0, t0 + steps_per_batch)

        def flat(a):
            a = a[sl]
            if layout == "env_major":
                a = np.swapaxes(a, 0, 1)
            return torch.from_numpy(np.ascontiguousarray(a).reshape(-1, *a.shape[2:])).to(dev)

        n = steps_per_batch * n_envs
        td = TensorDict({"pixels": flat(obs), "action": flat(act), ("next", "pixels"): flat(nxt),
                         ("next", "reward"): flat(rew), ("next", "done"): flat(done[..., None]),
                         "is_init": flat(init[..., None])}, [n])
        yield td


def _drive(storage_kwargs, n_envs, total_steps, steps_per_batch, layout, *, seed, pad, use_init, dev, max_size,
           min_len=1, max_len=40):
    """Writes the stream through a round-robin cursor into the storage and into the materialised oracle; checks every
    stored transition after every write, and the frame log against the restatement."""
    st = FrameStackStorage(max_size, n_envs=n_envs, batch_layout=layout, device=dev, **storage_kwargs)
    orc = fo.MaterialisedStorage(max_size)
    pool = head = last = None
    cursor = 0
    for td in _batches(n_envs, total_steps, steps_per_batch, layout, seed=seed, pad=pad, dev=dev, min_len=min_len,
                       max_len=max_len):
        n = td.shape[0]
        if not use_init:
            td.pop("is_init")
        slots = (cursor + np.arange(n)) % max_size
        o, x = td.get("pixels").cpu().numpy(), td.get(("next", "pixels")).cpu().numpy()
        st.set(torch.from_numpy(slots).to(dev), td)
        orc.set(slots, o, x)
        cursor = (cursor + n) % max_size
        # the frame log against the plain-loop restatement
        if pool is None:
            pool = np.zeros(tuple(st._pool.shape), dtype=np.uint8)
            head, last = np.zeros(n_envs, dtype=np.int64), np.ones(n_envs, dtype=np.uint8)
        words = fo.push(pool, head, last, o, x, td.get("is_init").cpu().numpy().reshape(-1) if use_init else None,
                        td.get(("next", "done")).cpu().numpy().reshape(-1), n_envs=n_envs,
                        layout=0 if layout == "env_major" else 1, k=K, ring=st._ring)
        got_words = st._inner._leaves[[tuple(k) if not isinstance(k, str) else (k,) for k in st._inner._spec[1]]
                                      .index(("_frame_word",))]
        np.testing.assert_array_equal(got_words.cpu().numpy()[slots], words)
        np.testing.assert_array_equal(st._head.cpu().numpy(), head)
        # every stored transition, in a shuffled order, equals what the reference's storage would return
        index = np.random.default_rng(cursor).permutation(len(st))
        got = st.get(torch.from_numpy(index).to(dev))
        want_o, want_x = orc.get(index)
        _, _, evicted = fo.rebuild(pool, head, got_words.cpu().numpy()[index], k=K, ring=st._ring)
        assert not evicted.any()
        np.testing.assert_array_equal(got.get("pixels").cpu().numpy(), want_o)
        np.testing.assert_array_equal(got.get(("next", "pixels")).cpu().numpy(), want_x)
        assert got.get("action").shape == (len(index), 1)
    st.check_index_status()
    return st, orc


CASES = [  # n_envs, total_steps, steps_per_batch, layout, pad, use_init, max_size
    (1, 300, 20, "env_major", "same", True, 128),
    (1, 300, 30, "env_major", "constant", False, 96),
    (4, 160, 8, "env_major", "same", False, 256),
    (4, 160, 8, "time_major", "constant", True, 256),
    (3, 90, 1, "time_major", "same", False, 60),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"E{c[0]}-{c[3]}-{c[4]}-{'init' if c[5] else 'done'}")
@pytest.mark.parametrize("materialize", [True, False])
def test_framestack_equals_materialised_storage_host_logic(emul, case, materialize):
    n_envs, total, per, layout, pad, use_init, max_size = case
    st, _ = _drive({"materialize": materialize, "min_episode_length": 1, "validate": True}, n_envs, total, per, layout,
                   seed=3, pad=pad, use_init=use_init, dev="cpu", max_size=max_size)
    # the point of it: ~ (1 + (k + 1) / episode_length) frames per transition instead of 2 k
    assert st.frame_bytes_per_transition < 2 * K * 30 * (1 + K + 1)


def test_framestack_contract_host_logic(emul, tmp_path):
    """Through the replay buffer (round-robin writer, fused range write), int / slice reads, errors, checkpoint."""
    rb = TensorDictReplayBuffer(storage=FrameStackStorage(64, n_envs=2, device="cpu", min_episode_length=1),
                                sampler=RandomSampler(), batch_size=16)
    orc = fo.MaterialisedStorage(64)
    cursor = 0
    for td in _batches(2, 100, 5, "env_major", seed=9, pad="same"):
        n = td.shape[0]
        slots = (cursor + np.arange(n)) % 64
        orc.set(slots, td.get("pixels").numpy(), td.get(("next", "pixels")).numpy())
        idx = rb.extend(td)
        np.testing.assert_array_equal(idx.numpy(), slots)
        cursor = (cursor + n) % 64
    assert len(rb) == 64
    batch = rb.sample()
    ix = batch.get("index").numpy()
    np.testing.assert_array_equal(batch.get("pixels").numpy(), orc.get(ix)[0])
    np.testing.assert_array_equal(batch.get(("next", "pixels")).numpy(), orc.get(ix)[1])
    st = rb._storage
    np.testing.assert_array_equal(st[5].get("pixels").numpy(), orc.get(5)[0])
    np.testing.assert_array_equal(st[3:9].get(("next", "pixels")).numpy(), orc.get(np.arange(3, 9))[1])
    with pytest.raises(RuntimeError, match="stream order"):
        st[3] = st[4]
    with pytest.raises(KeyError):
        st.set(torch.arange(2), TensorDict({"pixels": torch.zeros(2, K, 6, 5)}, [2]))
    with pytest.raises(RuntimeError, match="divide"):
        st.set(torch.arange(3), next(_batches(1, 3, 3, "env_major", seed=1, pad="same")))
    # checkpoint round trip into a fresh storage
    st.dumps(tmp_path / "fs")
    st2 = FrameStackStorage(64, n_envs=2, device="cpu")
    st2.loads(tmp_path / "fs")
    every = torch.arange(64)
    a, b = st.get(every), st2.get(every)
    for key in ("pixels", ("next", "pixels"), "action", ("next", "reward")):
        assert torch.equal(a.get(key), b.get(key))
    st3 = FrameStackStorage(64, n_envs=2, device="cpu")
    st3.load_state_dict(st.state_dict())
    assert torch.equal(st3.get(every).get("pixels"), a.get("pixels"))


def test_framestack_eviction_is_reported_host_logic(emul):
    """Episodes of one step log k + 1 frames per transition: a pool sized for long episodes laps itself, and sampling an
    affected transition raises instead of returning somebody else's frames."""
    st = FrameStackStorage(64, device="cpu", frame_capacity=80)
    for td in _batches(1, 64, 8, "env_major", seed=2, pad="same", min_len=1, max_len=1):
        st.set(torch.arange(8) + len(st), td)
    st.get(torch.arange(50, 64))
    st.check_index_status()                       # the newest transitions are intact
    st.get(torch.arange(0, 8))
    with pytest.raises(RuntimeError, match="frame_capacity"):
        st.check_index_status()


def test_framestack_validate_rejects_non_stacks(emul):
    st = FrameStackStorage(32, device="cpu", validate=True)
    td = next(_batches(1, 8, 8, "env_major", seed=4, pad="same", min_len=20, max_len=30))
    td.set("pixels", torch.randint(0, 255, td.get("pixels").shape, dtype=torch.uint8))   # not the predecessor's next stack
    with pytest.raises(RuntimeError, match="frame stack"):
        st.set(torch.arange(8), td)


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"E{c[0]}-{c[3]}-{c[4]}-{'init' if c[5] else 'done'}")
@pytest.mark.parametrize("materialize", [True, False])
def test_framestack_equals_materialised_storage(case, materialize):
    n_envs, total, per, layout, pad, use_init, max_size = case
    _drive({"materialize": materialize, "min_episode_length": 1}, n_envs, total, per, layout, seed=3, pad=pad,
           use_init=use_init, dev="cuda", max_size=max_size)


@pytest.mark.gpu
def test_framestack_atari_shapes_prioritized_buffer():
    """84x84 frames (bulk-DMA rows), 8 envs, prioritized sampling + fused range write: batches equal the materialised
    LazyTensorStorage buffer fed the same stream and the same generator seed, bit for bit."""
    dev = torch.device("cuda")
    cap, E = 4096, 8

    def make(storage):
        return TensorDictPrioritizedReplayBuffer(alpha=0.6, beta=0.4, storage=storage, batch_size=256,
                                                 generator=torch.Generator(device=dev).manual_seed(5))

    a = make(FrameStackStorage(cap, n_envs=E, device=dev))
    b = make(LazyTensorStorage(cap, device=dev))
    for td in _batches(E, 768, 32, "env_major", seed=11, pad="same", frame=(84, 84), dev=dev, min_len=30, max_len=200):
        td.set("td_error", torch.rand(td.shape[0], device=dev))
        ia, ib = a.extend(td.clone()), b.extend(td.clone())
        assert torch.equal(ia, ib)
        for _ in range(2):
            x, y = a.sample(), b.sample()
            assert torch.equal(x.get("index"), y.get("index"))
            for key in ("pixels", ("next", "pixels"), "action", ("next", "reward"), "priority_weight"):
                assert torch.equal(x.get(key), y.get(key)), key
            pr = torch.rand(256, device=dev)
            a.update_priority(x.get("index"), pr)
            b.update_priority(y.get("index"), pr)
    a._storage.check_index_status()
    st = a._storage
    assert st.frame_bytes_per_transition < 1.3 * 84 * 84          # vs 8 * 84 * 84 materialised
    assert ops.backend().name != "oracle-emulator"


@pytest.mark.gpu
def test_framestack_eviction_is_reported():
    dev = "cuda"
    st = FrameStackStorage(64, device=dev, frame_capacity=80)
    for td in _batches(1, 64, 8, "env_major", seed=2, pad="same", min_len=1, max_len=1, dev=dev):
        st.set(torch.arange(8, device=dev) + len(st), td)
    st.get(torch.arange(50, 64, device=dev))
    st.check_index_status()
    st.get(torch.arange(0, 8, device=dev))
    with pytest.raises(RuntimeError, match="frame_capacity"):
        st.check_index_status()


def test_sharded_buffer_over_framestack_storage_host_logic(emul):
    """The sharded buffer with a FrameStackStorage shard: the exchanged row carries the k + 1 distinct frames (one window),
    and the batch equals the plain sharded buffer's fed the same stream with the same generator."""
    from rl_b200.data.sharded import ShardedPrioritizedReplayBuffer

    mk = lambda: torch.Generator().manual_seed(5)
    a = ShardedPrioritizedReplayBuffer(alpha=0.6, beta=0.4, capacity=128, batch_size=32, device="cpu", generator=mk(),
                                       storage=FrameStackStorage(128, n_envs=2, device="cpu", min_episode_length=1))
    b = ShardedPrioritizedReplayBuffer(alpha=0.6, beta=0.4, capacity=128, batch_size=32, device="cpu", generator=mk())
    for td in _batches(2, 80, 8, "env_major", seed=21, pad="same"):
        td.set("td_error", torch.rand(td.shape[0], generator=torch.Generator().manual_seed(int(td.shape[0]))))
        a.extend(td.clone())
        b.extend(td.clone())
        x, y = a.sample(), b.sample()
        for key in ("index", "pixels", ("next", "pixels"), "action", ("next", "reward"), "priority_weight"):
            assert torch.equal(x.get(key), y.get(key)), key
        a.update_priority(x.get("index"), torch.ones(32))
        b.update_priority(y.get("index"), torch.ones(32))
    frame = 6 * 5
    assert b._layout.row - a._layout.row >= 3 * frame - 16      # 5 frames instead of 8 on the wire
    win = x.get("pixels")
    assert win.stride(0) >= 5 * frame                            # obs / next are views of one [B, k + 1, ...] window


def _slice_buffers(dev):
    from rl_b200.data import SliceSampler

    def make(storage):
        g = torch.Generator(device=dev).manual_seed(31)
        return TensorDictReplayBuffer(storage=storage, batch_size=64, generator=g,
                                      sampler=SliceSampler(num_slices=8, end_key=("next", "done"), strict_length=True))

    return make(FrameStackStorage(512, device=dev, min_episode_length=1)), make(LazyTensorStorage(512, device=dev))


def _check_slices(dev):
    """Trajectory slices (SliceSampler reads its episode signal through the storage's cheap view, the frames through the
    rebuilding gather): identical batches from the de-duplicated and the materialised storage."""
    a, b = _slice_buffers(dev)
    for td in _batches(1, 600, 60, "env_major", seed=17, pad="constant", dev=dev, min_len=9, max_len=40):
        a.extend(td.clone())
        b.extend(td.clone())
        x, y = a.sample(), b.sample()
        for key in ("index", "pixels", ("next", "pixels"), "action", ("next", "done"), ("next", "truncated")):
            assert torch.equal(x.get(key), y.get(key)), key
    a._storage.check_index_status()


def test_slice_sampler_over_framestack_storage_host_logic(emul):
    _check_slices("cpu")


@pytest.mark.gpu
def test_slice_sampler_over_framestack_storage():
    _check_slices("cuda")
