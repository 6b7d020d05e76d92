"""Kernel timeline (start / end per kernel, rank 0) of the graph-replayed N > 1 bench step via torch.profiler (CUPTI).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/timeline_dist.py [rows_per_shard]
"""
import os
import sys
from collections import defaultdict
from pathlib import Path

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from rl_b200 import ops  # noqa: E402
from rl_b200.graphs import CudaGraphStep  # noqa: E402

rank, world, local = bench.dist_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
be = ops.backend()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
rb, g = bench.build_sharded(dev, N, world, rank)
ring = bench.gae_ring(dev, bench.GAE_ROWS, bench.GAE_T, g, min_bytes=8 * bench.GAE_ROWS * bench.GAE_T * 14)
ring8 = [(v, nv, r, d.view(torch.uint8), t.view(torch.uint8)) for v, nv, r, d, t in ring]
td_loc = torch.rand(bench.BATCH, device=dev, generator=g)
rb.record_index_event = True
side_gae, side_upd = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
rb.sample_now()


def make_step(slot):
    v, nv, r, d8, t8 = ring8[slot]

    def step():
        main = torch.cuda.current_stream(dev)
        side_gae.wait_stream(main)
        with torch.cuda.stream(side_gae):
            a, tg = be.gae(v, nv, r, d8, t8, 0.99, 0.9405, bench.GAE_ROWS, bench.GAE_T, 1)
        batch = rb.sample(slot=slot)
        side_upd.wait_event(rb.index_ready)
        with torch.cuda.stream(side_upd):
            rb.update_local_priority(td_loc)
        main.wait_stream(side_upd)
        main.wait_stream(side_gae)
        return batch, a, tg

    return step


R = bench.N_BUFFERS


def group():
    outs = [make_step(k)() for k in range(R)]
    rb.join_exchange()
    return outs


graphs = [CudaGraphStep(group, generators=[rb.sampler._rng], warmup=1)]
for i in range(6):
    graphs[0]()
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(2):
        graphs[0]()
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    for e in evs:
        print(f"{e.time_range.start - t0:9.2f} -> {e.time_range.end - t0:9.2f} us  ({e.time_range.end - e.time_range.start:6.2f})  {e.name[:70]}")
    agg = defaultdict(list)
    for e in evs:
        agg[e.name[:50]].append(e.time_range.end - e.time_range.start)
    for k, v in agg.items():
        print(f"AVG {sum(v)/len(v):7.2f} us x{len(v):2d}  {k}")
    print(f"SPAN {evs[-1].time_range.end - evs[0].time_range.start:.1f} us for {2 * R} steps, world {world}")
dist.barrier()
dist.destroy_process_group()
