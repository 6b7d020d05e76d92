"""A few hot-path steps (sample -> update_priority -> GAE) through the public API, for the ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        -k regex:'gather_kernel|per_sample|tree_update|per_update|upd_|gae_|distribution' python profiles/prof_step.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from rl_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
be = ops.backend()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rb, g = bench.build_buffer(dev, N, seed=0)
ring = bench.gae_ring(dev, bench.GAE_ROWS, bench.GAE_T, g)
td_err = torch.rand(bench.BATCH, device=dev, generator=g)
torch.cuda.synchronize()
print("STEPS_BEGIN", flush=True)
for i in range(steps):
    v, nv, r, d, t = ring[i % len(ring)]
    batch = rb.sample()
    rb.update_priority(batch.get("index"), td_err)
    a, tg = be.gae(v, nv, r, d.view(torch.uint8), t.view(torch.uint8), 0.99, 0.9405, bench.GAE_ROWS, bench.GAE_T, 1)
torch.cuda.synchronize()
print("done")
