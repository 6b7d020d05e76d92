"""Kernel timeline (start/end, stream) of the CUDA-graph-replayed bench step via torch.profiler (CUPTI)."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from rl_b200 import ops  # noqa: E402
from rl_b200.graphs import CudaGraphStep  # noqa: E402

dev = torch.device("cuda", 0)
be = ops.backend()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
multi = (sys.argv[2] != "seq") if len(sys.argv) > 2 else True
rb, g = bench.build_buffer(dev, N, seed=0)
ring = bench.gae_ring(dev, bench.GAE_ROWS, bench.GAE_T, g)
v, nv, r, d, t = ring[0]
d8, t8 = d.view(torch.uint8), t.view(torch.uint8)
td_err = torch.rand(bench.BATCH, device=dev, generator=g)
smp = rb.sampler
smp.record_index_event = True
side_gae, side_upd = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
import os
if os.environ.get('RLB_PIN_L2') == '1':
    print('pin_l2 ->', smp.pin_l2(), smp.pin_l2(side_upd))


def step():
    main = torch.cuda.current_stream(dev)
    if multi:
        side_gae.wait_stream(main)
        with torch.cuda.stream(side_gae):
            a, tg = be.gae(v, nv, r, d8, t8, 0.99, 0.9405, bench.GAE_ROWS, bench.GAE_T, 1)
        batch = rb.sample()
        side_upd.wait_event(smp.index_ready)
        with torch.cuda.stream(side_upd):
            rb.update_priority(batch.get("index"), td_err)
        main.wait_stream(side_upd)
        main.wait_stream(side_gae)
    else:
        batch = rb.sample()
        rb.update_priority(batch.get("index"), td_err)
        a, tg = be.gae(v, nv, r, d8, t8, 0.99, 0.9405, bench.GAE_ROWS, bench.GAE_T, 1)
    return batch, a, tg


def three():
    return [step() for _ in range(3)]


gs = CudaGraphStep(three, generators=[g], warmup=2)
for _ in range(5):
    gs()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    gs()
    gs()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    print(f"{e.time_range.start - t0:9.2f} -> {e.time_range.end - t0:9.2f} us  ({e.time_range.end - e.time_range.start:6.2f})  {e.name[:70]}")
from collections import defaultdict
agg = defaultdict(list)
for e in evs:
    agg[e.name[:50]].append(e.time_range.end - e.time_range.start)
for k, v in agg.items():
    print(f"AVG {sum(v)/len(v):7.2f} us x{len(v):2d}  {k}")
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"SPAN {span:.1f} us for 6 steps")
