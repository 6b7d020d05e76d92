"""world_size-2 (gloo, CPU) tests of the capacity-sharded buffer's host logic: per-rank draws are index-exact
against the reference sampler glue on that shard, the gathered batch is the rank-order concatenation, weights
are identical on every rank, and priority write-back only touches the owning shard."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from _emul import OracleBackend
        from oracle import per_oracle as po
        from rl_b200 import ops
        from rl_b200.data import TensorDict
        from rl_b200.data.sharded import ShardedPrioritizedReplayBuffer

        ops.set_backend(OracleBackend())
        cap, B, alpha, beta = 1000, 64, 0.6, 0.4
        g = torch.Generator().manual_seed(100 + rank)
        rb = ShardedPrioritizedReplayBuffer(alpha=alpha, beta=beta, capacity=cap, batch_size=B, device="cpu",
                                            generator=g)
        assert rb.shard_capacity == 500 and rb.world == world
        gd = torch.Generator().manual_seed(7 + rank)
        n = 300 + 50 * rank
        data = TensorDict({"obs": torch.randn(n, 3, generator=gd), "frame": torch.randint(0, 255, (n, 2, 4), dtype=torch.uint8, generator=gd),
                           "flag": torch.rand(n, 1, generator=gd) < 0.5, "td_error": torch.rand(n, generator=gd) * 2}, [n])
        gidx = rb.extend(data)
        assert gidx.tolist() == list(range(rank * 500, rank * 500 + n))
        # reference glue on this shard
        ref = po.OraclePrioritizedSampler(500, alpha, beta)
        ref.mark_update(torch.arange(n))
        ref.update_priority(torch.arange(n), data.get("td_error"))
        state = g.get_state()
        batch = rb.sample()
        g2 = torch.Generator()
        g2.set_state(state)
        want_local, _ = ref.sample(n, B // world, generator=g2)
        assert torch.equal(rb.local_index, want_local)                       # index-exact per shard
        mine = {"index": want_local + rank * 500, "obs": data.get("obs")[want_local],
                "frame": data.get("frame")[want_local], "flag": data.get("flag")[want_local],
                "p": torch.as_tensor(ref._sum_tree[want_local.numpy()]),
                "S": torch.tensor(ref._sum_tree.query(0, n)), "m": torch.tensor(ref._min_tree.query(0, n))}
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        for key in ("index", "obs", "frame", "flag"):                          # rank-order concatenation
            assert torch.equal(batch.get(key), torch.cat([e[key] for e in everyone])), key
        assert batch.get("frame").dtype == torch.uint8 and batch.get("flag").dtype == torch.bool
        p = torch.cat([e["p"] for e in everyone])
        S = torch.cat([e["S"].expand(B // world) for e in everyone])
        mn = min(float(e["m"] / e["S"]) for e in everyone)
        want_w = torch.pow((p / S) / mn, -beta)
        torch.testing.assert_close(batch.get("priority_weight"), want_w, rtol=1e-6, atol=0)
        ws = [None] * world
        dist.all_gather_object(ws, batch.get("priority_weight"))
        assert all(torch.equal(ws[0], w) for w in ws)                          # identical on every rank
        # write-back with GLOBAL indices: only the owner changes
        before = rb.sampler._sum_tree.dump_leaves().clone()
        new_p = torch.full((B,), 5.0)
        rb.update_priority(batch.get("index"), new_p)
        after = rb.sampler._sum_tree.dump_leaves()
        own = (batch.get("index") // 500) == rank
        touched = torch.zeros(500, dtype=torch.bool)
        touched[(batch.get("index")[own] - rank * 500)] = True
        assert torch.equal(after[~touched], before[~touched])
        assert torch.allclose(after[touched], torch.tensor(5.0 + 1e-8) ** alpha)
        # pipelined (sample() returns the previous draw) against plain, same seeds, no write-backs in between: the same
        # sequence of global batches on both ranks; flush() hands out the draw still in flight
        bufs = []
        for pipe in (True, False):
            gp = torch.Generator().manual_seed(200 + rank)
            rbp = ShardedPrioritizedReplayBuffer(alpha=alpha, beta=beta, capacity=cap, batch_size=B, device="cpu",
                                                 generator=gp, pipeline=pipe)
            rbp.extend(data.clone())
            bufs.append(rbp)
        for it in range(4):
            x = bufs[0].sample() if it < 3 else bufs[0].flush()
            y = bufs[1].sample()
            for key in ("index", "obs", "frame", "flag", "priority_weight"):
                assert torch.equal(x.get(key), y.get(key)), ("pipelined", it, key)
        assert bufs[0].flush() is None
        # a FrameStackStorage shard next to a plain one: same global batches, k + 1 frames per exchanged row
        from rl_b200.data import FrameStackStorage
        from test_framestack import _batches

        pair = []
        for fs in (True, False):
            gp = torch.Generator().manual_seed(300 + rank)
            pair.append(ShardedPrioritizedReplayBuffer(
                alpha=alpha, beta=beta, capacity=256, batch_size=16, device="cpu", generator=gp,
                storage=FrameStackStorage(128, n_envs=2, device="cpu", min_episode_length=1) if fs else None))
        for td in _batches(2, 48, 8, "env_major", seed=60 + rank, pad="constant"):
            td.set("td_error", torch.rand(td.shape[0], generator=torch.Generator().manual_seed(4)))
            for rbp in pair:
                rbp.extend(td.clone())
            x, y = (rbp.sample() for rbp in pair)
            for key in ("index", "pixels", ("next", "pixels"), "action", ("next", "reward"), "priority_weight"):
                assert torch.equal(x.get(key), y.get(key)), ("framestack shard", key)
            for rbp, batch in zip(pair, (x, y)):
                rbp.update_priority(batch.get("index"), torch.full((16,), 0.5 + rank))
        assert pair[0]._layout.row < pair[1]._layout.row
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))


def test_sharded_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"


def test_sharded_single_process(emul):
    """World size 1: the sharded buffer degenerates to the plain prioritized buffer (weights = (p/p_min)^-beta)."""
    from rl_b200.data import LazyTensorStorage, TensorDict, TensorDictPrioritizedReplayBuffer
    from rl_b200.data.sharded import ShardedPrioritizedReplayBuffer

    mk = lambda: torch.Generator().manual_seed(5)
    data = TensorDict({"x": torch.randn(200, 5), "td_error": torch.rand(200)}, [200])
    a = ShardedPrioritizedReplayBuffer(alpha=0.6, beta=0.4, capacity=256, batch_size=32, device="cpu", generator=mk())
    b = TensorDictPrioritizedReplayBuffer(alpha=0.6, beta=0.4, storage=LazyTensorStorage(256, device="cpu"),
                                          batch_size=32, generator=mk())
    a.extend(data.clone())
    b.extend(data.clone())
    sa, sb = a.sample(), b.sample()
    assert torch.equal(sa.get("index"), sb.get("index")) and torch.equal(sa.get("x"), sb.get("x"))
    torch.testing.assert_close(sa.get("priority_weight"), sb.get("priority_weight"), rtol=1e-6, atol=0)


def test_sharded_pipelined_returns_previous_draw(emul):
    """pipeline=True: sample() issues draw k and returns draw k-1, so without write-backs in between the returned
    sequence is the plain buffer's; flush() hands out the draw still in flight; update_priority on the returned index
    vector (or a clone of it) rewrites this shard's rows only."""
    from rl_b200.data import TensorDict
    from rl_b200.data.sharded import ShardedPrioritizedReplayBuffer

    mk = lambda: torch.Generator().manual_seed(11)
    data = TensorDict({"x": torch.randn(300, 5), "td_error": torch.rand(300)}, [300])
    a = ShardedPrioritizedReplayBuffer(alpha=0.6, beta=0.4, capacity=512, batch_size=16, device="cpu", generator=mk(),
                                       pipeline=True)
    b = ShardedPrioritizedReplayBuffer(alpha=0.6, beta=0.4, capacity=512, batch_size=16, device="cpu", generator=mk())
    assert a.n_buffers == 4 and b.n_buffers == 2
    a.extend(data.clone())
    b.extend(data.clone())
    want = [b.sample() for _ in range(4)]
    got = [a.sample() for _ in range(3)]
    got.append(a.flush())
    assert a.flush() is None
    for w, g in zip(want, got):
        assert torch.equal(w.get("index"), g.get("index")) and torch.equal(w.get("x"), g.get("x"))
        torch.testing.assert_close(w.get("priority_weight"), g.get("priority_weight"), rtol=1e-6, atol=0)
    # the fast path keys on the storage of the returned index tensor, not on object identity; a clone takes the
    # general path and must give the same trees
    last = got[-1].get("index")
    pr = torch.rand(16) + 1.0
    a.update_priority(last, pr)
    b.update_priority(want[-1].get("index").clone(), pr)
    assert torch.equal(a.sampler._sum_tree.values, b.sampler._sum_tree.values)
    # update_local_priority: the rows of the latest local draw
    c = ShardedPrioritizedReplayBuffer(alpha=0.6, beta=0.4, capacity=512, batch_size=16, device="cpu", generator=mk())
    c.extend(data.clone())
    sc = c.sample()
    c.update_local_priority(pr)
    d = ShardedPrioritizedReplayBuffer(alpha=0.6, beta=0.4, capacity=512, batch_size=16, device="cpu", generator=mk())
    d.extend(data.clone())
    sd = d.sample()
    d.update_priority(sd.get("index"), pr)
    assert torch.equal(sc.get("index"), sd.get("index"))
    assert torch.equal(c.sampler._sum_tree.values, d.sampler._sum_tree.values)
