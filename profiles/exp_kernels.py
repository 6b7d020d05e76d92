"""In-situ kernel timing: the step pipeline [rand, per_sample, gather, update, gae] replayed from a CUDA graph, and
leave-one-out variants, so each kernel's contribution is measured in its real cache/TLB context (not ncu's
cold serialised replay).  Prints microseconds per step."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from rl_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
be = ops.backend()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rb, g = bench.build_buffer(dev, N, seed=0)
smp, st = rb.sampler, rb.storage
ring = bench.gae_ring(dev, bench.GAE_ROWS, bench.GAE_T, g)
ring8 = [(v, nv, r, d.view(torch.uint8), t.view(torch.uint8)) for v, nv, r, d, t in ring]
td_err = torch.rand(B, device=dev, generator=g)
static_idx = torch.randint(0, N, (B,), device=dev, generator=g)
static_u = torch.rand(B, device=dev, generator=g)
plan = be.gather_plan(st._leaves)
R = len(ring)


def make(parts):
    def fn():
        outs = []
        for slot in range(R):
            u = torch.rand(B, device=dev, generator=g) if "rand" in parts else static_u
            if "sample" in parts:
                idx, w = be.per_sample(smp._sum_tree.values, smp._min_tree.values, smp._max_capacity,
                                       smp._sum_tree.capacity, N, u, 0.4, True)
            else:
                idx = static_idx
            if "gather" in parts:
                outs.append(plan.run(idx, N))
            if "update" in parts:
                be.per_update(smp._sum_tree.values, smp._min_tree.values, smp._sum_tree.capacity, idx, td_err, 0.6, 1e-8,
                              smp._max_priority_buf, smp._tree_workspace(B), 1)
            if "gae" in parts:
                v, nv, r, d8, t8 = ring8[slot]
                outs.append(be.gae(v, nv, r, d8, t8, 0.99, 0.9405, bench.GAE_ROWS, bench.GAE_T, 1))
        return outs
    return fn


def time_graph(parts, reps=20):
    stream = torch.cuda.Stream(dev)
    graph = torch.cuda.CUDAGraph()
    graph.register_generator_state(g)
    fn = make(parts)
    with torch.cuda.stream(stream):
        fn()
        stream.synchronize()
        with torch.cuda.graph(graph, stream=stream):
            keep = fn()
        graph.replay()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            graph.replay()
        e1.record(stream)
        stream.synchronize()
    del keep
    return e0.elapsed_time(e1) * 1e3 / (reps * R)


ALL = ["rand", "sample", "gather", "update", "gae"]
full = time_graph(ALL)
print(f"full pipeline           {full:8.2f} us/step")
for p in ALL:
    t = time_graph([q for q in ALL if q != p])
    print(f"  without {p:8s}      {t:8.2f} us/step   (delta {full - t:6.2f})")
for p in ALL:
    print(f"  only {p:8s}         {time_graph([p]):8.2f} us/step")
