"""The priority write-back kernel alone (B items on a 1M-slot tree pair), for ncu source-level profiles and ticks."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rl_b200.data import PrioritizedSampler  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
g = torch.Generator(device=dev).manual_seed(0)
smp = PrioritizedSampler(N, 0.6, 0.4, device=dev)
smp.update_priority(torch.arange(N, device=dev), torch.rand(N, device=dev, generator=g))
torch.cuda.synchronize()
for it in range(6):
    idx = torch.randint(0, N, (B,), device=dev, generator=g)
    smp.update_priority(idx, torch.rand(B, device=dev, generator=g))
torch.cuda.synchronize()
# graph-replayed timing of 20 back-to-back updates
from bench import graph_us  # noqa: E402

idxs = [torch.randint(0, N, (B,), device=dev, generator=g) for _ in range(20)]
pr = torch.rand(B, device=dev, generator=g)
print(f"update_priority B={B} N={N}: {graph_us([(lambda ix=ix: smp.update_priority(ix, pr)) for ix in idxs], dev):.2f} us per call (graph-replayed)")
