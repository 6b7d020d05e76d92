"""Two-or-more-GPU check of the sharded buffer (run under torchrun on a GPU box):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_check.py

Every rank fills its shard, then for both transports ("nvlink": gather kernel broadcasts through NVLink peer memory
+ signal-pad barrier; "nccl": all-gather) the gathered batch must be the rank-order concatenation of every rank's
index-exact local draw, identical on all ranks, and the two transports must agree bit for bit."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import per_oracle as po  # noqa: E402
from rl_b200.data import TensorDict  # noqa: E402
from rl_b200.data.sharded import ShardedPrioritizedReplayBuffer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cap, B = 40_000 * world, 128 * world
    n = 30_000
    gd = torch.Generator(device=dev).manual_seed(50 + rank)
    data = TensorDict({"pixels": torch.randint(0, 255, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=gd),
                       "action": torch.randint(0, 18, (n, 1), device=dev, generator=gd),
                       "vec": torch.randn(n, 17, device=dev, generator=gd),
                       "flag": torch.rand(n, 1, device=dev, generator=gd) < 0.5,
                       "td_error": torch.rand(n, device=dev, generator=gd)}, [n])
    results = {}
    for transport in ("nvlink", "nccl"):
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        rb = ShardedPrioritizedReplayBuffer(alpha=0.6, beta=0.4, capacity=cap, batch_size=B, device=dev, generator=g,
                                            transport=transport)
        rb.extend(data.clone())
        outs = []
        for it in range(4):  # exercises both halves of the double buffer twice
            smp = rb.sampler
            leaves = smp._sum_tree.dump_leaves().cpu().numpy()
            os_, om = po.OracleTree(rb.shard_capacity, False), po.OracleTree(rb.shard_capacity, True)
            os_.load_leaves(leaves)
            om.load_leaves(smp._min_tree.dump_leaves().cpu().numpy())
            state = g.get_state()
            batch = rb.sample()
            torch.cuda.synchronize()
            g2 = torch.Generator(device=dev)
            g2.set_state(state)
            u = torch.rand(B // world, device=dev, generator=g2)
            want_local, _, _, _ = po.per_sample_c(os_, om, n, u.cpu().numpy(), 0.4)
            np.testing.assert_array_equal(rb.local_index.cpu().numpy(), want_local)          # index-exact per shard
            wl = torch.from_numpy(want_local).to(dev)
            mine = {"index": (wl + rank * rb.shard_capacity).cpu(), "pixels": data.get("pixels")[wl].cpu(),
                    "vec": data.get("vec")[wl].cpu(), "flag": data.get("flag")[wl].cpu(),
                    "action": data.get("action")[wl].cpu()}
            everyone = [None] * world
            dist.all_gather_object(everyone, mine)
            for k in ("index", "pixels", "vec", "flag", "action"):
                assert torch.equal(batch.get(k).cpu(), torch.cat([e[k] for e in everyone])), (transport, it, k)
            ws = [None] * world
            dist.all_gather_object(ws, batch.get("priority_weight").cpu())
            assert all(torch.equal(ws[0], w) for w in ws)
            outs.append({k: batch.get(k).clone() for k in ("index", "pixels", "priority_weight")})
            rb.update_priority(batch.get("index"), torch.rand(B, device=dev, generator=torch.Generator(device=dev).manual_seed(it)))
        results[transport] = outs
        assert (rb._symm not in (None, False)) == (transport == "nvlink")
    for a, b in zip(results["nvlink"], results["nccl"]):
        for k in a:
            assert torch.equal(a[k], b[k]), k
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check ok: world={world}, nvlink == nccl == rank-order concat of index-exact local draws")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
