"""The TorchRL-facing API on a real GPU: full buffer flows against the reference glue, CUDA-graph replay,
the sharded buffer at world size 1, BASELINE-sized configurations through size-independent properties."""
import numpy as np
import pytest
import torch

from oracle import per_oracle as po

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def _oracle_from(smp, filled):
    """Oracle trees holding exactly the device trees' leaf bits."""
    N = smp._max_capacity
    leaves = smp._sum_tree.dump_leaves().cpu().numpy()
    os_, om = po.OracleTree(N, False), po.OracleTree(N, True)
    os_.load_leaves(leaves)
    ml = smp._min_tree.dump_leaves().cpu().numpy()
    om.load_leaves(ml)
    return os_, om


def test_c2_sized_prioritized_buffer_index_exact(cuda_backend):
    """BASELINE configs[1] at full tree size (1M capacity, B=256, alpha=.6, beta=.4); rows are kept narrow so the
    test is quick -- the gather itself is size-checked in test_gpu_kernels.  Sampled indices are bit-exact against
    the C oracle on identical leaves for the draws of the buffer's own CUDA generator; the returned rows are the
    storage rows at those indices; after the TD-error write-back the whole heap equals the serial reference."""
    from rl_b200.data import LazyTensorStorage, TensorDict, TensorDictPrioritizedReplayBuffer

    N, B = 1_000_000, 256
    g = torch.Generator(device=dev()).manual_seed(0)
    rb = TensorDictPrioritizedReplayBuffer(alpha=0.6, beta=0.4, storage=LazyTensorStorage(N, device=dev()),
                                           batch_size=B, generator=g)
    for lo in range(0, N, 250_000):
        n = 250_000
        rb.extend(TensorDict({"obs": torch.randint(0, 255, (n, 64), dtype=torch.uint8, device=dev(), generator=g),
                              "reward": torch.randn(n, device=dev(), generator=g),
                              "td_error": torch.rand(n, device=dev(), generator=g)}, [n]))
    assert len(rb) == N
    smp = rb.sampler
    os_, om = _oracle_from(smp, N)
    np.testing.assert_array_equal(smp._sum_tree.values.cpu().numpy()[1:], os_.values()[1:])
    np.testing.assert_array_equal(smp._min_tree.values.cpu().numpy()[1:], om.values()[1:])
    for _ in range(3):
        state = g.get_state()
        batch = rb.sample()
        g.set_state(state)
        u = torch.rand(B, device=dev(), generator=g)
        want_idx, want_w, _, _ = po.per_sample_c(os_, om, N, u.cpu().numpy(), 0.4)
        idx = batch.get("index")
        np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
        np.testing.assert_allclose(batch.get("priority_weight").cpu().numpy(), want_w, rtol=2e-6)
        full = rb.storage.get(slice(None))
        assert torch.equal(batch.get("obs"), full.get("obs")[idx])
        assert torch.equal(batch.get("reward"), full.get("reward")[idx])
        td = torch.rand(B, device=dev(), generator=g)
        batch.set("td_error", td)
        rb.update_tensordict_priority(batch)
        leaf = torch.pow(td + 1e-8, 0.6).cpu().numpy()   # the device's own pow bits (fused == torch.pow is tested)
        os_[want_idx] = leaf
        om[want_idx] = leaf
        np.testing.assert_array_equal(smp._sum_tree.values.cpu().numpy()[1:], os_.values()[1:])
        np.testing.assert_array_equal(smp._min_tree.values.cpu().numpy()[1:], om.values()[1:])


def test_cuda_graph_step_replays_fresh_draws(cuda_backend):
    from rl_b200.data import LazyTensorStorage, TensorDict, TensorDictPrioritizedReplayBuffer
    from rl_b200.graphs import CudaGraphStep
    from rl_b200.objectives.value import vec_generalized_advantage_estimate

    N, B = 50_000, 128
    g = torch.Generator(device=dev()).manual_seed(1)
    rb = TensorDictPrioritizedReplayBuffer(alpha=0.7, beta=0.5, storage=LazyTensorStorage(N, device=dev()),
                                           batch_size=B, generator=g)
    data = TensorDict({"x": torch.randn(N, 33, device=dev(), generator=g),
                       "td_error": torch.rand(N, device=dev(), generator=g)}, [N])
    rb.extend(data)
    td = torch.rand(B, device=dev(), generator=g)
    v, nv, r = (torch.randn(64, 128, 1, device=dev(), generator=g) for _ in range(3))
    done = torch.rand(64, 128, 1, device=dev(), generator=g) < 0.05

    def step():
        batch = rb.sample()
        rb.update_priority(batch.get("index"), td)
        return batch, vec_generalized_advantage_estimate(0.99, 0.95, v, nv, r, done=done)

    graphed = CudaGraphStep(step, generators=[g], warmup=2)
    seen = []
    os_, om = None, None
    for it in range(4):
        os_, om = _oracle_from(rb.sampler, N)
        state = g.get_state()
        batch, (adv, tgt) = graphed()
        torch.cuda.synchronize()
        idx = batch.get("index").clone()
        # the replay drew what an eager call would have drawn from the same generator state
        g2 = torch.Generator(device=dev())
        g2.set_state(state)
        u = torch.rand(B, device=dev(), generator=g2)
        want_idx, _, _, _ = po.per_sample_c(os_, om, N, u.cpu().numpy(), 0.5)
        np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
        assert torch.equal(batch.get("x"), data.get("x")[idx])
        seen.append(idx)
    assert not torch.equal(seen[0], seen[1])  # fresh random draws on every replay
    fa, _ = po.gae_f64(0.99, 0.95, v.cpu(), nv.cpu(), r.cpu(), done.cpu(), done.cpu())
    torch.testing.assert_close(adv.cpu().double(), fa, rtol=1e-5, atol=1e-5)


def test_sharded_world1_equals_plain_buffer(cuda_backend):
    from rl_b200.data import LazyTensorStorage, TensorDict, TensorDictPrioritizedReplayBuffer
    from rl_b200.data.sharded import ShardedPrioritizedReplayBuffer

    N, B = 20_000, 256
    mk = lambda: torch.Generator(device=dev()).manual_seed(9)
    gd = torch.Generator(device=dev()).manual_seed(2)
    data = TensorDict({"pixels": torch.randint(0, 255, (N, 4, 84, 84), dtype=torch.uint8, device=dev(), generator=gd),
                       "action": torch.randint(0, 18, (N, 1), device=dev(), generator=gd),
                       "flag": torch.rand(N, 1, device=dev(), generator=gd) < 0.5,
                       "vec": torch.randn(N, 17, device=dev(), generator=gd),
                       "td_error": torch.rand(N, device=dev(), generator=gd)}, [N])
    a = ShardedPrioritizedReplayBuffer(alpha=0.6, beta=0.4, capacity=N, batch_size=B, device=dev(), generator=mk())
    b = TensorDictPrioritizedReplayBuffer(alpha=0.6, beta=0.4, storage=LazyTensorStorage(N, device=dev()),
                                          batch_size=B, generator=mk())
    a.extend(data.clone())
    b.extend(data.clone())
    for _ in range(2):
        sa, sb = a.sample(), b.sample()
        assert torch.equal(sa.get("index"), sb.get("index"))
        for k in ("pixels", "action", "flag", "vec"):
            assert torch.equal(sa.get(k), sb.get(k)), k          # strided views into the packed buffer
        torch.testing.assert_close(sa.get("priority_weight"), sb.get("priority_weight"), rtol=1e-6, atol=0)
        td = torch.rand(B, device=dev(), generator=gd)
        a.update_priority(sa.get("index"), td)
        b.update_priority(sb.get("index"), td)
    assert torch.equal(a.sampler._sum_tree.values[1:], b.sampler._sum_tree.values[1:])


def test_l2_persist_and_index_event(cuda_backend):
    from rl_b200.data import LazyTensorStorage, PrioritizedSampler, ReplayBuffer

    smp = PrioritizedSampler(10_000, 0.6, 0.4, device=dev())
    assert smp.pin_l2() >= 0
    assert cuda_backend.l2_persist(None) == 0
    rb = ReplayBuffer(storage=LazyTensorStorage(10_000, device=dev()), sampler=smp, batch_size=64)
    rb.extend(torch.arange(5000.0, device=dev()))
    smp.record_index_event = True
    side = torch.cuda.Stream(dev())
    out, info = rb.sample(return_info=True)
    side.wait_event(smp.index_ready)
    with torch.cuda.stream(side):
        rb.update_priority(info["index"], torch.ones(64, device=dev()))
    torch.cuda.synchronize()
    assert torch.equal(out, info["index"].float())


def test_index_check_mode(cuda_backend):
    from rl_b200.data import TensorStorage

    st = TensorStorage(torch.arange(100.0, device=dev()).reshape(50, 2), device=dev())
    st.get(torch.tensor([0, 49], device=dev()))            # the status word is on by default for tensor indices
    st.check_index_status()
    last = st.get(torch.tensor([49], device=dev()))[0].clone()
    out = st.get(torch.tensor([0, 50], device=dev()))      # out of range: reads the clamped row ...
    assert torch.equal(out[1], last)
    with pytest.raises(IndexError):
        st.check_index_status()                            # ... and the synchronising check says so
    # deferred form: no sync anywhere -- a LATER tensor-indexed call of the same storage raises
    st.get(torch.tensor([0, 77], device=dev()))
    torch.cuda.synchronize()
    with pytest.raises(IndexError, match="earlier"):
        st.get(torch.tensor([1], device=dev()))
    # an out-of-range WRITE is dropped, never redirected onto another slot
    before = st.get(slice(None)).clone()
    st.set(torch.tensor([3, 50, -51], device=dev()), torch.full((3, 2), -1.0, device=dev()))
    after = st.get(slice(None))
    assert torch.equal(after[3], torch.full((2,), -1.0, device=dev()))
    keep = torch.ones(50, dtype=torch.bool, device=dev())
    keep[3] = False
    assert torch.equal(after[keep], before[keep])
    with pytest.raises(IndexError):
        st.check_index_status()
    st.enable_index_check(False)
    st.get(torch.tensor([0, 99], device=dev()))
    st.check_index_status()


def test_c5_shaped_gather_and_writeback(cuda_backend):
    """BASELINE configs[4] shapes (obs 376 f32, act 17 f32, B=4096) on one shard: narrow / 4-byte-aligned rows go
    through the vector role; TD-error write-back of 4096 priorities uses the stamp-dedupe general path."""
    from rl_b200.data import LazyTensorStorage, TensorDict, TensorDictPrioritizedReplayBuffer

    N, B = 200_000, 4096
    g = torch.Generator(device=dev()).manual_seed(5)
    rb = TensorDictPrioritizedReplayBuffer(alpha=0.6, beta=0.4, storage=LazyTensorStorage(N, device=dev()),
                                           batch_size=B, generator=g)
    data = TensorDict({"obs": torch.randn(N, 376, device=dev(), generator=g),
                       "action": torch.randn(N, 17, device=dev(), generator=g),
                       "next": {"obs": torch.randn(N, 376, device=dev(), generator=g),
                                "reward": torch.randn(N, device=dev(), generator=g),
                                "done": torch.rand(N, 1, device=dev(), generator=g) < 0.01}}, [N])
    rb.extend(data)
    batch = rb.sample()
    idx = batch.get("index")
    for k in ("obs", "action", ("next", "obs"), ("next", "reward"), ("next", "done")):
        assert torch.equal(batch.get(k), data.get(k)[idx]), k
    os_, om = _oracle_from(rb.sampler, N)
    td = torch.rand(B, device=dev(), generator=g)
    rb.update_priority(idx, td)
    leaf = torch.pow(td + 1e-8, 0.6).cpu().numpy()
    os_[idx.cpu().numpy()] = leaf      # serial, input order, last duplicate wins
    om[idx.cpu().numpy()] = leaf
    np.testing.assert_array_equal(rb.sampler._sum_tree.values.cpu().numpy()[1:], os_.values()[1:])
    np.testing.assert_array_equal(rb.sampler._min_tree.values.cpu().numpy()[1:], om.values()[1:])


@pytest.mark.parametrize("with_td_error", [False, True])
def test_fused_write_path_equals_general_path_and_oracle(cuda_backend, monkeypatch, with_td_error):
    """rb.extend through rlb_extend (rows + default priorities, one launch) leaves storage, both heaps
    and the running max bit-identical to the general path (scatter / slice copies + sorted-merge update) and to the
    restated reference sampler, over many batches that wrap around the ring, with write-backs in between."""
    from rl_b200.data import LazyTensorStorage, TensorDict, TensorDictPrioritizedReplayBuffer
    from rl_b200.data.writers import RoundRobinWriter

    N = 3000
    g = torch.Generator(device=dev()).manual_seed(7)

    def make():
        return TensorDictPrioritizedReplayBuffer(alpha=0.7, beta=0.5, storage=LazyTensorStorage(N, device=dev()),
                                                 batch_size=64, generator=torch.Generator(device=dev()).manual_seed(1))

    fused, general = make(), make()
    orc = po.OraclePrioritizedSampler(N, 0.7, 0.5)
    launches = []
    real_extend = cuda_backend.extend
    monkeypatch.setattr(cuda_backend, "extend", lambda *a, **k: (launches.append(1), real_extend(*a, **k))[1])
    rng = np.random.default_rng(0)
    cursor = 0
    for it in range(14):
        n = int(rng.integers(1, 1500))
        data = {"pixels": torch.randint(0, 255, (n, 2, 84, 84), dtype=torch.uint8, device=dev(), generator=g),
                "action": torch.randint(0, 6, (n, 1), device=dev(), generator=g),
                "reward": torch.randn(n, device=dev(), generator=g),
                "done": torch.rand(n, 1, device=dev(), generator=g) < 0.1}
        if with_td_error:
            data["td_error"] = torch.rand(n, device=dev(), generator=g) * (it + 1)
        before = len(launches)
        idx_f = fused.extend(TensorDict(dict(data), [n]))
        assert len(launches) == before + 1                        # the whole write is one rlb_extend launch
        with monkeypatch.context() as m:
            m.setattr(RoundRobinWriter, "_extend_fused", lambda self, *a: False)
            idx_g = general.extend(TensorDict(dict(data), [n]))
        assert len(launches) == before + 1
        want = (cursor + torch.arange(n)) % N
        assert torch.equal(idx_f.cpu(), want) and torch.equal(idx_g.cpu(), want)
        orc.mark_update(want)
        if with_td_error:
            orc.update_priority(want, data["td_error"].cpu())
        cursor = (cursor + n) % N
        assert len(fused) == len(general)
        for a, b in zip(fused.storage._leaves, general.storage._leaves):
            assert torch.equal(a[:len(fused)], b[:len(fused)])      # (slots never written are uninitialised memory)
        # exact against the general path and against the reference tree rebuilt from the same leaf bits; the CPU
        # sampler restatement agrees to an ulp (its leaves come from glibc powf, the device's from CUDA powf -- the
        # function the reference's own CUDA path calls)
        os_, om = _oracle_from(fused.sampler, N)
        for tree, ot in (("_sum_tree", os_), ("_min_tree", om)):
            got = getattr(fused.sampler, tree).values.cpu().numpy()
            np.testing.assert_array_equal(got, getattr(general.sampler, tree).values.cpu().numpy())
            np.testing.assert_array_equal(got[1:], ot.values()[1:])
            cap = fused.sampler._sum_tree.capacity
            np.testing.assert_allclose(got[cap:cap + N], getattr(orc, tree).values()[cap:cap + N], rtol=1e-6)
        assert torch.equal(fused.sampler._max_priority_buf, general.sampler._max_priority_buf)
        if it % 4 == 3:
            k = 100
            ix = torch.from_numpy(rng.integers(0, len(fused), k)).to(dev())
            pr = torch.rand(k, device=dev(), generator=g) * 5
            for rb in (fused, general):
                rb.update_priority(ix, pr)
            orc.update_priority(ix.cpu(), pr.cpu())
    a, b = fused.sample(), general.sample()
    assert torch.equal(a.get("index"), b.get("index")) and torch.equal(a.get("pixels"), b.get("pixels"))
    # the row in the storage is the row that was written
    assert torch.equal(fused.storage.get(slice(None)).get("reward")[idx_f], data["reward"])


def test_extend_rows_only_and_non_prioritized_buffers(cuda_backend):
    """Buffers without trees (RandomSampler) use the same launch with no tree role; more leaves than RLB_MAX_LEAVES
    split into groups; odd row widths take the vector role."""
    from rl_b200.data import LazyTensorStorage, TensorDict, TensorDictReplayBuffer

    N = 257
    rb = TensorDictReplayBuffer(storage=LazyTensorStorage(N, device=dev()), batch_size=8)
    g = torch.Generator(device=dev()).manual_seed(3)
    ref = {}
    cursor = 0
    for it in range(5):
        n = [100, 200, 57, 257, 1][it]
        data = {f"k{j}": torch.randn(n, j % 5 + 1, device=dev(), generator=g) for j in range(30)}
        data["bytes"] = torch.randint(0, 255, (n, 4099), dtype=torch.uint8, device=dev(), generator=g)
        data["wide"] = torch.randn(n, 2048, device=dev(), generator=g)
        rb.extend(TensorDict(dict(data), [n]))
        ix = (cursor + torch.arange(n)) % N
        cursor = (cursor + n) % N
        for k, v in data.items():
            ref.setdefault(k, torch.zeros(N, *v.shape[1:], dtype=v.dtype, device=dev()))[ix.to(dev())] = v
        full = rb.storage.get(slice(None))
        for k in data:
            assert torch.equal(full.get(k), ref[k][:len(rb)]), (it, k)


@pytest.mark.parametrize("strict", [True, False])
def test_slice_sampler_buffer_on_device(cuda_backend, strict):
    """TensorDictReplayBuffer + SliceSampler on the GPU: the slices are the oracle's for the draws of the buffer's own
    CUDA generator, rows come from the sampled slots, done = stored done | truncated; cache_values survives sampling and
    is dropped by a write; a wrapped ring keeps trajectories that cross the end of the storage."""
    from oracle import slice_oracle as so
    from rl_b200.data import LazyTensorStorage, SliceSampler, TensorDict, TensorDictReplayBuffer

    L, S, T = 50_000, 32, 16
    g = torch.Generator(device=dev()).manual_seed(5)
    rb = TensorDictReplayBuffer(storage=LazyTensorStorage(L, device=dev()), batch_size=S * T, generator=g,
                                sampler=SliceSampler(num_slices=S, end_key=("next", "done"), strict_length=strict,
                                                     cache_values=True))
    rng = np.random.default_rng(1)
    written = 0
    for n in (30_000, 15_000, 12_000):                # the third batch wraps around
        done = torch.from_numpy(rng.random(n) < 0.04).reshape(n, 1)
        rb.extend(TensorDict({"obs": torch.arange(written, written + n, dtype=torch.float32, device=dev()).reshape(n, 1),
                              ("next", "done"): done.to(dev())}, [n]))
        written += n
        full = rb.storage.get(slice(None))
        stored_done = full.get(("next", "done")).reshape(-1).cpu().numpy()
        filled = len(rb)
        cursor = rb.storage._last_cursor
        cursor = cursor.stop - 1 if isinstance(cursor, slice) else int(cursor.reshape(-1)[-1])
        start, stop, length = so.traj_table(end=stored_done, at_capacity=filled == L, cursor=cursor)
        variable = (not strict) and bool((length < T).any())
        vs, _, vl = so.valid_trajectories(start, stop, length, T, strict)
        for _ in range(2):
            state = g.get_state()
            batch = rb.sample()
            g.set_state(state)
            traj = torch.randint(len(vs), (S,), device=dev(), generator=g)
            u = torch.rand(S, device=dev(), generator=g)
            oi, otr, _, _ = so.slice_index(vs, vl, seq_length=T, num_slices=S, storage_length=L,
                                           traj_draw=traj.cpu().numpy(), u=u.cpu().numpy(), strict_length=strict)
            idx = batch.get("index").reshape(-1)
            np.testing.assert_array_equal(idx.cpu().numpy(), oi)
            assert variable == (len(oi) != S * T)
            np.testing.assert_array_equal(batch.get(("next", "truncated")).reshape(-1).cpu().numpy(), otr)
            np.testing.assert_array_equal(batch.get(("next", "done")).reshape(-1).cpu().numpy(), stored_done[oi] | otr)
            assert torch.equal(batch.get("obs"), full.get("obs")[idx])
        assert rb.sampler._cache
    assert rb.storage._is_full


@pytest.mark.parametrize("strict", [True, False])
def test_prioritized_slice_sampler_buffer_on_device(cuda_backend, strict):
    """TensorDictReplayBuffer + PrioritizedSliceSampler on the GPU at a 1M-slot ring (strict_length=False: slices stop
    with their trajectory, the batch shrinks): starts are the oracle's for the
    draws of the buffer's CUDA generator and the device's own leaves; slices stay inside one trajectory; every step
    carries its start's weight; the true priorities are untouched by sampling."""
    from oracle import slice_oracle as so
    from rl_b200.data import LazyTensorStorage, PrioritizedSliceSampler, TensorDict, TensorDictReplayBuffer

    L, S, T = 1_000_000, 64, 32
    g = torch.Generator(device=dev()).manual_seed(9)
    rb = TensorDictReplayBuffer(storage=LazyTensorStorage(L, device=dev()), batch_size=S * T, generator=g,
                                sampler=PrioritizedSliceSampler(L, 0.6, 0.4, num_slices=S, end_key=("next", "done"),
                                                                cache_values=True, strict_length=strict))
    rng = np.random.default_rng(2)
    for n in (600_000, 400_000, 123_456):             # fills the ring, then wraps
        done = torch.from_numpy(rng.random(n) < (0.01 if strict else 0.04)).reshape(n, 1).to(dev())
        rb.extend(TensorDict({"t": torch.arange(n, device=dev()).reshape(n, 1), ("next", "done"): done}, [n]))
        filled = len(rb)
        ix = torch.from_numpy(rng.integers(0, filled, 50_000)).to(dev())
        rb.update_priority(ix, torch.rand(50_000, device=dev(), generator=g) * 4)
        before = rb.sampler._sum_tree.values.clone()
        os_, om = _oracle_from(rb.sampler, L)
        orc = po.OraclePrioritizedSampler(L, 0.6, 0.4)
        orc._sum_tree, orc._min_tree = os_, om
        stored_done = rb.storage.get(slice(None)).get(("next", "done")).reshape(-1).cpu().numpy()
        cursor = rb.storage._last_cursor
        cursor = cursor.stop - 1 if isinstance(cursor, slice) else int(cursor.reshape(-1)[-1])
        start, stop, length = so.traj_table(end=stored_done, at_capacity=filled == L, cursor=cursor)
        for _ in range(2):
            state = g.get_state()
            batch = rb.sample()
            g.set_state(state)
            u = torch.rand(S, device=dev(), generator=g)
            oi, ow, otr, _ = so.prioritized_slice_sample(orc, start, stop, length, seq_length=T, num_slices=S,
                                                         storage_len=filled, u=u.cpu().numpy(), strict_length=strict)
            idx = batch.get("index").reshape(-1).cpu().numpy()
            np.testing.assert_array_equal(idx, oi)
            np.testing.assert_allclose(batch.get("priority_weight").reshape(-1).cpu().numpy(), ow, rtol=2e-6)
            np.testing.assert_array_equal(batch.get(("next", "truncated")).reshape(-1).cpu().numpy(), otr)
            if strict:
                assert not stored_done[idx.reshape(S, T)[:, :-1]].any()          # no slice crosses a trajectory end
                w = batch.get("priority_weight").reshape(S, T)
                assert (w == w[:, :1]).all()
            else:
                assert len(idx) < S * T                                          # some slice was cut at its trajectory's end
                assert not stored_done[idx[~otr]].any()
        assert torch.equal(rb.sampler._sum_tree.values, before)


def test_slice_sampler_without_replacement_on_device(cuda_backend):
    """A sweep of SliceSamplerWithoutReplacement on the GPU visits every stored trajectory exactly once (slices lie in
    distinct trajectories until ran_out), slices are consecutive steps, and shuffle=False walks the ring in order."""
    from oracle import slice_oracle as so
    from rl_b200.data import LazyTensorStorage, SliceSamplerWithoutReplacement, TensorDict, TensorDictReplayBuffer

    L, S, T = 20_000, 8, 4
    rng = np.random.default_rng(4)
    done = rng.random(L) < 0.02
    done[::37] = False
    length = so.traj_table(end=done, at_capacity=True, cursor=L - 1)[2]
    for shuffle in (True, False):
        rb = TensorDictReplayBuffer(storage=LazyTensorStorage(L, device=dev()), batch_size=S * T,
                                    generator=torch.Generator(device=dev()).manual_seed(0),
                                    sampler=SliceSamplerWithoutReplacement(num_slices=S, end_key=("next", "done"),
                                                                           shuffle=shuffle, strict_length=False))
        rb.extend(TensorDict({"t": torch.arange(L, device=dev()).reshape(L, 1),
                              ("next", "done"): torch.from_numpy(done).reshape(L, 1).to(dev())}, [L]))
        start, stop, _ = so.traj_table(end=done, at_capacity=True, cursor=L - 1)
        owner = np.empty(L, dtype=np.int64)                       # slot -> trajectory id
        for k, (s0, e0) in enumerate(zip(start, stop)):
            if s0 <= e0:
                owner[s0:e0 + 1] = k
            else:
                owner[s0:] = k
                owner[:e0 + 1] = k
        seen = []
        for _ in range(10_000):
            b = rb.sample()
            first = b.get("index").reshape(-1)[b.get(("next", "truncated")).reshape(-1).roll(1)].cpu().numpy()
            seen.extend(owner[first].tolist())
            if rb.sampler.ran_out:
                break
        assert sorted(seen) == list(range(len(start)))           # each trajectory exactly once per sweep
        if not shuffle:
            assert seen == list(range(len(start)))
    assert (length > 0).all()


def test_predraw_same_stream_eager_and_graph(cuda_backend):
    """PrioritizedSampler.predraw: (1) eager, the sampled indices are those of the plain sampler for the same seed;
    (2) captured, every replay consumes what the previous replay drew -- index-exact against the oracle for the uniforms
    the generator produced one step earlier -- and replays keep producing fresh batches."""
    from rl_b200.data import LazyTensorStorage, TensorDict, TensorDictPrioritizedReplayBuffer
    from rl_b200.graphs import CudaGraphStep

    N, B = 20_000, 64

    def make(predraw, seed=4):
        g = torch.Generator(device=dev()).manual_seed(seed)
        rb = TensorDictPrioritizedReplayBuffer(alpha=0.7, beta=0.5, storage=LazyTensorStorage(N, device=dev()),
                                               batch_size=B, generator=g)
        rb.sampler.predraw = predraw
        dg = torch.Generator(device=dev()).manual_seed(11)
        rb.extend(TensorDict({"x": torch.randn(N, 8, device=dev(), generator=dg),
                              "td_error": torch.rand(N, device=dev(), generator=dg)}, [N]))
        return rb, g

    plain, _ = make(False)
    early, _ = make(True)
    td = torch.rand(B, device=dev(), generator=torch.Generator(device=dev()).manual_seed(2))
    for _ in range(5):
        a, b = plain.sample(), early.sample()
        assert torch.equal(a.get("index"), b.get("index"))
        plain.update_priority(a.get("index"), td)
        early.update_priority(b.get("index"), td)

    rb, g = make(True, seed=9)

    def step():
        batch = rb.sample()
        rb.update_priority(batch.get("index"), td)
        return batch

    graphed = CudaGraphStep(step, generators=[g], warmup=2)
    seen = []
    for it in range(4):
        os_, om = _oracle_from(rb.sampler, N)
        u = rb.sampler._u_next.clone()                   # drawn by the previous call / replay
        batch = graphed()
        torch.cuda.synchronize()
        idx = batch.get("index").clone()
        want_idx, _, _, _ = po.per_sample_c(os_, om, N, u.cpu().numpy(), 0.5)
        np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
        assert not torch.equal(rb.sampler._u_next, u)    # and the replay refilled the buffer for the next one
        seen.append(idx)
    assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[2])


def test_storage_checkpoint_roundtrip_on_device(cuda_backend, tmp_path):
    """TensorStorageCheckpointer from / to HBM: filled rows stream through a pinned buffer into the tensordict-memmap
    layout and back (checkpointers.py:326-455); buffer.dumps / loads carries storage + sampler + writer."""
    import json

    from rl_b200.data import LazyTensorStorage, TensorDict, TensorDictPrioritizedReplayBuffer

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    mk = lambda: TensorDictPrioritizedReplayBuffer(alpha=0.6, beta=0.4, storage=LazyTensorStorage(5000, device=dev),
                                                   batch_size=64, generator=torch.Generator(device=dev).manual_seed(5))
    rb = mk()
    data = TensorDict({"pixels": torch.randint(0, 255, (3000, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                       "next": {"reward": torch.randn(3000, device=dev, generator=g),
                                "done": torch.rand(3000, 1, device=dev, generator=g) < 0.1},
                       "td_error": torch.rand(3000, device=dev, generator=g)}, [3000])
    rb.extend(data)
    rb.dumps(tmp_path / "ckpt")
    meta = json.loads((tmp_path / "ckpt" / "storage" / "storage_metadata.json").read_text())
    assert meta["len"] == 3000 and meta["is_pytree"] is False
    assert (tmp_path / "ckpt" / "storage" / "pixels.memmap").stat().st_size == 5000 * 4 * 84 * 84
    rb2 = mk()
    rb2.extend(data[:10])          # initialised with something else
    rb2.loads(tmp_path / "ckpt")
    assert len(rb2) == 3000
    idx = torch.arange(3000, device=dev)
    a, b = rb.storage.get(idx), rb2.storage.get(idx)
    for k in ("pixels", ("next", "reward"), ("next", "done")):
        assert torch.equal(a.get(k), b.get(k)), k
    assert torch.equal(rb.sampler._sum_tree.values, rb2.sampler._sum_tree.values)
    s1, s2 = rb.sample(), rb2.sample()
    assert torch.equal(s1.get("index"), s2.get("index")) and torch.equal(s1.get("pixels"), s2.get("pixels"))


@pytest.mark.parametrize("mode", ["strict", "loose", "span"])
def test_slice_sampler_2d_storage_on_device(cuda_backend, mode):
    """SliceSampler on an ndim=2 storage ([T, E]: one ring per column) on the GPU: (time, column) index pairs, flags
    and rows against the oracle's ring-by-ring table for the draws of the buffer's own CUDA generator (the same
    construction is pinned to the unmodified reference on CPU, tests/test_host_logic.py)."""
    from oracle import slice_oracle as so
    from rl_b200.data import LazyTensorStorage, SliceSampler, TensorDict, TensorDictReplayBuffer

    dev = torch.device("cuda", 0)
    Tm, E, S, T = 6000, 8, 24, 12
    kw = dict(strict=dict(), loose=dict(strict_length=False), span=dict(span=(True, 3)))[mode]
    g = torch.Generator(device=dev).manual_seed(9)
    rb = TensorDictReplayBuffer(storage=LazyTensorStorage(Tm * E, device=dev, ndim=2), batch_size=S * T, generator=g,
                                sampler=SliceSampler(num_slices=S, end_key=("next", "done"), **kw), dim_extend=0)
    rng = np.random.default_rng(4)
    written = 0
    for n in (4000, 1500, 1700):                      # the third batch wraps around
        done = torch.from_numpy(rng.random((n, E, 1)) < 0.05)
        obs = (torch.arange(written, written + n).view(n, 1, 1) * 16 + torch.arange(E).view(1, E, 1)).float()
        rb.extend(TensorDict({"obs": obs.to(dev), ("next", "done"): done.to(dev)}, [n, E]))
        written += n
        rows = min(Tm, written)
        stored = rb.storage._leaves
        full = rb.storage.get(slice(None))
        stored_done = full.get(("next", "done"))[:rows].reshape(rows, E).cpu().numpy()
        cursor = rb.storage._last_cursor
        cursor = cursor.stop - 1 if isinstance(cursor, slice) else int(torch.as_tensor(cursor).reshape(-1)[-1])
        st, sp, ln, col = so.traj_table_nd(end=stored_done, at_capacity=rows == Tm, cursor=cursor)
        strict = mode != "loose"
        if strict:
            keep = ln >= T
            st, sp, ln, col = st[keep], sp[keep], ln[keep], col[keep]
        for _ in range(2):
            state = g.get_state()
            batch = rb.sample()
            g.set_state(state)
            traj = torch.randint(len(st), (S,), device=dev, generator=g)
            u = torch.rand(S, device=dev, generator=g)
            span = (-1, 3) if mode == "span" else (0, 0)
            oi, otr, _, oseq = so.slice_index(st, ln, seq_length=T, num_slices=S, storage_length=rows, traj_draw=traj.cpu().numpy(),
                                              u=u.cpu().numpy(), strict_length=strict, span=span,
                                              force_variable=mode == "span")
            ocol = np.repeat(col[traj.cpu().numpy()], oseq)
            idx = batch.get("index")
            np.testing.assert_array_equal(idx[..., 0].reshape(-1).cpu().numpy(), oi)
            np.testing.assert_array_equal(idx[..., 1].reshape(-1).cpu().numpy(), ocol)
            np.testing.assert_array_equal(batch.get(("next", "truncated")).reshape(-1).cpu().numpy(), otr)
            np.testing.assert_array_equal(batch.get(("next", "done")).reshape(-1).cpu().numpy(), stored_done[oi, ocol] | otr)
            want_obs = full.get("obs")[torch.from_numpy(oi).to(dev), torch.from_numpy(ocol).to(dev)]
            assert torch.equal(batch.get("obs").reshape(-1), want_obs.reshape(-1))


def test_sampler_status_surfaces_without_sync(cuda_backend):
    """the CPU reference raises "non-positive p_sum" inside sample(); here the kernel sets a bit and a LATER
    sample() raises once the asynchronous host mirror of the status word has landed (check_status() raises at once)."""
    from rl_b200.data import PrioritizedSampler

    class _St:
        ndim, shape, device = 1, (100,), torch.device("cuda", 0)

        def __len__(self):
            return 100

    smp = PrioritizedSampler(100, 0.6, 0.4, device=torch.device("cuda", 0))
    smp._rng = torch.Generator(device="cuda").manual_seed(0)
    smp.status_check_every = 1
    smp.sample(_St(), 8)                 # the trees are empty: p_sum == 0
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="non-positive p_sum"):
        for _ in range(3):               # the copy armed by one call is seen by a later one
            smp.sample(_St(), 8)
            torch.cuda.synchronize()
    smp.update_priority(torch.arange(100, device="cuda"), torch.rand(100, device="cuda") + 0.1)
    smp.check_status() if False else smp._status.zero_()
    for _ in range(3):
        smp.sample(_St(), 8)
        torch.cuda.synchronize()
    smp.check_status()
