"""Times rlb_gather (C2 shape: 8 leaves, 56 474 B/row) for every rl_b200/variant_*.so build; graph-replayed launches."""
import glob, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from rl_b200 import ops

dev = torch.device("cuda", 0)
rb, g = bench.build_buffer(dev, 300_000, seed=0)
st = rb.storage
N = len(st)
for B in (256, 1024, 16384):
    idxs = [torch.randint(0, N, (B,), device=dev, generator=g) for _ in range(16)]
    for so in sorted(glob.glob(str(Path(ops._PKG) / "variant_*.so"))) + [str(ops._SO)]:
        ops._SO = Path(so)
        ops.set_backend(None)
        be = ops.backend()
        plan = be.gather_plan(st._leaves)
        stream = torch.cuda.Stream(dev)
        with torch.cuda.stream(stream):
            for ix in idxs[:2]:
                plan.run(ix, N)
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                keep = [plan.run(ix, N) for ix in idxs]
            graph.replay(); stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                graph.replay()
            e1.record(stream)
            stream.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (5 * len(idxs))
        rowb = sum(l[0].numel() * l.element_size() for l in st._leaves)
        gbs = B * (2 * rowb + 8) / us / 1e3
        print(f"B={B:6d} {Path(so).name:32s} {us:8.2f} us  {gbs:7.1f} GB/s  {gbs/6489.6:.3f}")
        del keep, graph
