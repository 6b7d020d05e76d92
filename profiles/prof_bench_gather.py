"""The gather / GAE / sample / update kernels at the BENCH configuration (8 leaves, 1M-row Atari buffer, B=256; GAE
[4096,128]), launched the way bench.py times them -- a CUDA graph of back-to-back launches on cold rows -- for

    ncu --set full --graph-profiling node --cache-control none --clock-control none --import-source on \
        -k regex:"gather_kernel|gae_rows|per_sample|tree_update" -o gpurun_out/prof_r2_bench python profiles/prof_bench_gather.py

(node-level graph profiling keeps the launches inside their graph; no cache flush between them, like the timed run)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from rl_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
be = ops.backend()
N = int(sys.argv[1]) if len(sys.argv) > 1 else bench.CAPACITY
rb, g = bench.build_buffer(dev, N, seed=0)
st = rb.storage
ring = bench.gae_ring(dev, bench.GAE_ROWS, bench.GAE_T, g)
n_launch = 12
idxs = [torch.randint(0, len(st), (bench.BATCH,), device=dev, generator=g) for _ in range(n_launch)]
td = torch.rand(bench.BATCH, device=dev, generator=g)
stream = torch.cuda.Stream(dev)
with torch.cuda.stream(stream):
    for ix in idxs[:2]:
        be.gather(st._leaves, ix, len(st))
    rb.sample()
    rb.update_priority(idxs[0], td)
    stream.synchronize()
    gr = torch.cuda.CUDAGraph()
    gr.register_generator_state(rb.sampler._rng)
    with torch.cuda.graph(gr, stream=stream):
        keep = [be.gather(st._leaves, ix, len(st)) for ix in idxs]
        keep += [be.gae(x[0], x[1], x[2], x[3].view(torch.uint8), x[4].view(torch.uint8), bench.GAMMA,
                        bench.GAMMA * bench.LMBDA, bench.GAE_ROWS, bench.GAE_T, 1) for x in ring[:n_launch]]
        for ix in idxs[:6]:
            keep.append(rb.sample())
            rb.update_priority(ix, td)
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        gr.replay()
        e1.record(stream)
        stream.synchronize()
print(f"graph of {n_launch} gathers + {n_launch} GAE + 6 x (sample, update): {e0.elapsed_time(e1) * 1e3:.1f} us per replay (not a bench figure when run under ncu)")
