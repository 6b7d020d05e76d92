"""The graph-replayed three-stream bench step (sample ‖ update ‖ GAE) and the gather alone, for every
rl_b200/variant_*.so build of the gather ring geometry (RLB_GATHER_PIPES / STAGES / AHEAD / CHUNK): how much of the
update kernel's slow-down inside the step comes from the DMA bytes the gather keeps in flight."""
import glob
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from rl_b200 import ops  # noqa: E402
from rl_b200.graphs import CudaGraphStep  # noqa: E402

dev = torch.device("cuda", 0)
rb, g = bench.build_buffer(dev, int(sys.argv[1]) if len(sys.argv) > 1 else 400_000, seed=0)
ring = bench.gae_ring(dev, bench.GAE_ROWS, bench.GAE_T, g)
td_err = torch.rand(bench.BATCH, device=dev, generator=g)
smp, st = rb.sampler, rb.storage
smp.record_index_event = True
side_gae, side_upd = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
SPG = 20


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for so in [str(ops._SO)] + sorted(glob.glob(str(Path(ops._PKG) / "variant_*.so"))):
    ops._SO = Path(so)
    ops.set_backend(None)
    be = ops.backend()
    st._plan = None

    def step(i):
        main = torch.cuda.current_stream(dev)
        v, nv, r, d, t = ring[i % len(ring)]
        side_gae.wait_stream(main)
        with torch.cuda.stream(side_gae):
            a, tg = be.gae(v, nv, r, d.view(torch.uint8), t.view(torch.uint8), 0.99, 0.9405, bench.GAE_ROWS, bench.GAE_T, 1)
        batch = rb.sample()
        side_upd.wait_event(smp.index_ready)
        with torch.cuda.stream(side_upd):
            rb.update_priority(batch.get("index"), td_err)
        main.wait_stream(side_upd)
        main.wait_stream(side_gae)
        return batch, a, tg

    gs = CudaGraphStep(lambda: [step(i) for i in range(SPG)], generators=[g], warmup=1)
    for _ in range(3):
        gs()
    us_step = min(timed(gs, 20) for _ in range(3)) / SPG
    idxs = [torch.randint(0, len(st), (bench.BATCH,), device=dev, generator=g) for _ in range(16)]
    us_gather = bench.graph_us([(lambda ix=ix: be.gather(st._leaves, ix, len(st))) for ix in idxs], dev)
    big = [torch.randint(0, len(st), (4096,), device=dev, generator=g) for _ in range(4)]
    us_big = bench.graph_us([(lambda ix=ix: be.gather(st._leaves, ix, len(st))) for ix in big], dev)
    us_upd = bench.graph_us([(lambda ix=ix: rb.update_priority(ix, td_err)) for ix in idxs], dev)
    print(f"{Path(so).name:36s} step {us_step:6.2f} us   gather B=256 alone {us_gather:6.2f} us   B=4096 {us_big:6.1f} us   "
          f"update B=256 alone {us_upd:6.2f} us", flush=True)
    del gs
