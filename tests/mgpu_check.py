"""Two-or-more-GPU check of the sharded buffer (run under torchrun on a GPU box; tests/test_mgpu.py does that when the
box has >= 2 GPUs):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_check.py

Every rank fills its shard, then for the three exchange modes -- "nvlink" (gather kernel broadcasts through NVLink peer
memory, flags close the exchange), "nvlink-pipelined" (sample() returns the previous draw) and "nccl" (all-gather) --
the gathered batch must be the rank-order concatenation of every rank's index-exact local draw (SURVEY.md section 8e
"Parity definition for W>1": rank r's draw is index-exact vs the oracle sampler on that shard's priorities for the
draws of its own generator), identical on all ranks, and the modes must agree bit for bit.  Also checked: a
CUDA-graph-captured pipelined step replayed cyclically over the receive slots, and the write-back paths."""
import os
import sys
from collections import deque
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import per_oracle as po  # noqa: E402
from rl_b200.data import TensorDict  # noqa: E402
from rl_b200.data.sharded import ShardedPrioritizedReplayBuffer  # noqa: E402

KEYS = ("index", "pixels", "vec", "flag", "action")


def expected_draws(rb, g, n, b_loc, n_draws, dev):
    """Oracle: the local indices of the next `n_draws` draws of generator `g` on the shard's current trees."""
    smp = rb.sampler
    os_, om = po.OracleTree(rb.shard_capacity, False), po.OracleTree(rb.shard_capacity, True)
    os_.load_leaves(smp._sum_tree.dump_leaves().cpu().numpy())
    om.load_leaves(smp._min_tree.dump_leaves().cpu().numpy())
    g2 = torch.Generator(device=dev)
    g2.set_state(g.get_state())
    out = []
    for _ in range(n_draws):
        u = torch.rand(b_loc, device=dev, generator=g2)
        want, _, _, _ = po.per_sample_c(os_, om, n, u.cpu().numpy(), 0.4)
        out.append(want)
    return out


def _mc_default():
    """Exercise the multicast path wherever the fabric offers it (the product's "auto" starts at 4 ranks)."""
    return {"0": False, "1": True}.get(os.environ.get("RLB_MGPU_MULTICAST", ""), "probe")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cap, B = 40_000 * world, 128 * world
    n = 30_000
    b_loc = B // world
    gd = torch.Generator(device=dev).manual_seed(50 + rank)
    data = TensorDict({"pixels": torch.randint(0, 255, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=gd),
                       "action": torch.randint(0, 18, (n, 1), device=dev, generator=gd),
                       "vec": torch.randn(n, 17, device=dev, generator=gd),
                       "flag": torch.rand(n, 1, device=dev, generator=gd) < 0.5,
                       "td_error": torch.rand(n, device=dev, generator=gd)}, [n])

    def gathered_expectation(want_local):
        wl = torch.from_numpy(want_local).to(dev)
        mine = {"index": (wl + rank * (cap // world)).cpu(), "pixels": data.get("pixels")[wl].cpu(),
                "vec": data.get("vec")[wl].cpu(), "flag": data.get("flag")[wl].cpu(),
                "action": data.get("action")[wl].cpu()}
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        return {k: torch.cat([e[k] for e in everyone]) for k in KEYS}

    results = {}
    mc_used = False
    for mode in ("nvlink", "nvlink-unicast", "nvlink-pipelined", "nccl"):
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        rb = ShardedPrioritizedReplayBuffer(alpha=0.6, beta=0.4, capacity=cap, batch_size=B, device=dev, generator=g,
                                            transport=mode.split("-")[0], pipeline=mode.endswith("pipelined"),
                                            multicast=False if mode.endswith("unicast") else _mc_default())
        rb.extend(data.clone())
        outs, queue = [], deque()
        for it in range(6):  # every receive slot comes round at least once
            n_new = 2 if (rb.pipeline and it == 0) else 1
            for want_local in expected_draws(rb, g, n, b_loc, n_new, dev):
                queue.append(gathered_expectation(want_local))
            batch = rb.sample()
            torch.cuda.synchronize()
            rb.check_exchange()
            want = queue.popleft()
            if not rb.pipeline:
                np.testing.assert_array_equal(rb.local_index.cpu().numpy() + rank * rb.shard_capacity,
                                              want["index"][rank * b_loc:(rank + 1) * b_loc].numpy())
            for k in KEYS:
                assert torch.equal(batch.get(k).cpu(), want[k]), (mode, it, k)
            ws = [None] * world
            dist.all_gather_object(ws, batch.get("priority_weight").cpu())
            assert all(torch.equal(ws[0], w) for w in ws), (mode, it, "weights differ between ranks")
            outs.append({k: batch.get(k).clone() for k in ("index", "pixels", "priority_weight")})
            # write-back of the returned batch (fast path: the returned index tensor) -- same priorities on every rank
            pr = torch.rand(B, device=dev, generator=torch.Generator(device=dev).manual_seed(it))
            if it % 2:
                rb.update_priority(batch.get("index"), pr)
            else:
                rb.update_priority(batch.get("index").clone(), pr)   # general path: global indices, kernel-side filter
        if rb.pipeline:
            last = rb.flush()
            torch.cuda.synchronize()
            want = queue.popleft()
            for k in KEYS:
                assert torch.equal(last.get(k).cpu(), want[k]), (mode, "flush", k)
        assert not queue
        results[mode] = outs
        assert (rb._symm not in (None, False)) == mode.startswith("nvlink")
        mc_used = mc_used or rb._mc_delta != 0
        assert not (mode.endswith("unicast") and rb._mc_delta)
    # the plain nvlink and nccl runs see identical trees and draws: bit-equal batches (the pipelined run applies its
    # write-backs one draw later, so its later batches legitimately differ)
    for name in ("nvlink", "nvlink-unicast"):
        for a, b in zip(results[name], results["nccl"]):
            for k in a:
                assert torch.equal(a[k], b[k]), (name, k)
    for k in results["nvlink"][0]:
        assert torch.equal(results["nvlink"][0][k], results["nvlink-pipelined"][0][k]), k

    # ---- a captured pipelined step, replayed cyclically over the slots, against eager expectations
    from rl_b200.graphs import CudaGraphStep

    g = torch.Generator(device=dev).manual_seed(99 + rank)
    rb = ShardedPrioritizedReplayBuffer(alpha=0.6, beta=0.4, capacity=cap, batch_size=B, device=dev, generator=g,
                                        transport="nvlink", pipeline=True)
    rb.extend(data.clone())
    rb.record_index_event = True
    pr_loc = torch.rand(b_loc, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    side = torch.cuda.Stream(dev)

    def make_step(slot):
        def step():
            main = torch.cuda.current_stream(dev)
            batch = rb.sample(slot=slot)
            side.wait_event(rb.index_ready)
            with torch.cuda.stream(side):
                rb.update_local_priority(pr_loc)   # needs the sampled indices only: overlaps the exchange
            main.wait_stream(side)
            rb.join_exchange()                     # the exchange stream must rejoin before the capture ends
            return batch, rb.local_index.clone()

        return step

    steps = [CudaGraphStep(make_step(i), generators=[g], warmup=1) for i in range(rb.n_buffers)]
    torch.cuda.synchronize()
    dist.barrier()
    prev_expect = None
    for it in range(3 * rb.n_buffers):
        want_local = expected_draws(rb, g, n, b_loc, 1, dev)[0]
        expect = gathered_expectation(want_local)
        batch, lidx = steps[it % rb.n_buffers]()
        torch.cuda.synchronize()
        rb.check_exchange()
        np.testing.assert_array_equal(lidx.cpu().numpy(), want_local)
        if prev_expect is not None:   # the step returns the draw of the previous replay
            for k in KEYS:
                assert torch.equal(batch.get(k).cpu(), prev_expect[k]), ("graph", it, k)
        prev_expect = expect
    # ---- a FrameStackStorage shard: k + 1 unique frames per transition on the wire, same batches as the plain shard
    from rl_b200.data import FrameStackStorage

    sys.path.insert(0, str(ROOT / "tests"))
    from test_framestack import _batches

    bufs = []
    for fs in (True, False):
        g = torch.Generator(device=dev).manual_seed(300 + rank)
        bufs.append(ShardedPrioritizedReplayBuffer(
            alpha=0.6, beta=0.4, capacity=4096 * world, batch_size=B, device=dev, generator=g, transport="nvlink",
            storage=FrameStackStorage(4096, n_envs=4, device=dev) if fs else None))
    for td in _batches(4, 256, 64, "env_major", seed=40 + rank, pad="same", frame=(84, 84), dev=dev, min_len=20, max_len=90):
        td.set("td_error", torch.rand(td.shape[0], device=dev, generator=torch.Generator(device=dev).manual_seed(3)))
        for rb in bufs:
            rb.extend(td.clone())
        x, y = (rb.sample() for rb in bufs)
        torch.cuda.synchronize()
        for rb in bufs:
            rb.check_exchange()
        for k in ("index", "pixels", ("next", "pixels"), "action", ("next", "reward"), "priority_weight"):
            assert torch.equal(x.get(k), y.get(k)), ("framestack shard", k)
        pr = torch.rand(B, device=dev, generator=torch.Generator(device=dev).manual_seed(9))
        for rb, batch in zip(bufs, (x, y)):
            rb.update_priority(batch.get("index"), pr)
    bufs[0].storage.check_index_status()
    assert bufs[0]._layout.row < 0.65 * bufs[1]._layout.row
    dist.barrier()
    if rank == 0:
        print(f"framestack shard: {bufs[0]._layout.row} B per exchanged transition instead of {bufs[1]._layout.row}")
        print(f"multicast (multimem.st through the NVSwitch) {'ACTIVE' if mc_used else 'not available: unicast copies'}")
        print(f"mgpu_check ok: world={world}, nvlink == nvlink-pipelined == nccl == rank-order concat of index-exact "
              f"local draws; captured pipelined step replayed over {rb.n_buffers} slots")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
