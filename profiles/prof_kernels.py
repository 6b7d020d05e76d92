"""Launches each hot-path kernel a few times at the benchmark shapes, for ncu.

    ncu --set full --clock-control none --import-source on -k regex:'gather_kernel|gae_rows|per_sample|tree_update' \
        -o gpurun_out/prof python profiles/prof_kernels.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rl_b200 import ops  # noqa: E402
from rl_b200.data import LazyTensorStorage, PrioritizedSampler, TensorDict  # noqa: E402

dev = torch.device("cuda", 0)
be = ops.backend()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000   # 11 GB of pixels: far larger than L2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
g = torch.Generator(device=dev).manual_seed(0)
st = LazyTensorStorage(N, device=dev)
td = TensorDict({"pixels": torch.randint(0, 256, (N, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                 "action": torch.randint(0, 18, (N, 1), device=dev, generator=g),
                 "next": {"pixels": torch.randint(0, 256, (N, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                          "reward": torch.randn(N, device=dev, generator=g),
                          "done": torch.rand(N, 1, device=dev, generator=g) < 0.01,
                          "terminated": torch.rand(N, 1, device=dev, generator=g) < 0.01}}, [N])
st.set(slice(0, N), td)
smp = PrioritizedSampler(1_000_000, 0.6, 0.4, device=dev)
smp.update_priority(torch.arange(1_000_000, device=dev), torch.rand(1_000_000, device=dev, generator=g))
v, nv, r = (torch.randn(4096, 128, 1, device=dev, generator=g) for _ in range(3))
term = torch.rand(4096, 128, 1, device=dev, generator=g) < 0.02
done = term | (torch.rand(4096, 128, 1, device=dev, generator=g) < 0.02)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for it in range(4):
    idx = torch.randint(0, N, (B,), device=dev, generator=g)
    flush.fill_(it)                                            # evict L2 between iterations
    for mode in (0, 1):
        out = be.gather(st._leaves, idx, N, mode=mode)
    u = torch.rand(B, device=dev, generator=g)
    i2, w = be.per_sample(smp._sum_tree.values, smp._min_tree.values, 1_000_000, smp._sum_tree.capacity, 1_000_000, u,
                          0.4, True)
    smp.update_priority(i2, torch.rand(B, device=dev, generator=g))
    flush.fill_(it + 1)
    a, t = be.gae(v, nv, r, done.view(torch.uint8), term.view(torch.uint8), 0.99, 0.9405, 4096, 128, 1)
torch.cuda.synchronize()
print("done")
