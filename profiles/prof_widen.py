"""Launches the SURVEY 8(f) kernels (write path, range update, scan siblings, trajectory table, slice expansion) for ncu.

    ncu --set full --clock-control none --import-source on -k regex:'extend_kernel|tree_range|traj_|slice_index|gae_' \
        -o gpurun_out/prof_widen python profiles/prof_widen.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rl_b200 import ops  # noqa: E402
from rl_b200.data import (LazyTensorStorage, PrioritizedSampler, TensorDict,  # noqa: E402
                          TensorDictPrioritizedReplayBuffer)
from rl_b200.objectives.value import vec_td_lambda_return_estimate, vtrace_advantage_estimate  # noqa: E402

dev = torch.device("cuda", 0)
be = ops.backend()
g = torch.Generator(device=dev).manual_seed(0)
N, n = 200_000, 1024
rb = TensorDictPrioritizedReplayBuffer(alpha=0.6, beta=0.4, storage=LazyTensorStorage(N, device=dev), batch_size=256)
data = TensorDict({"pixels": torch.randint(0, 255, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                   ("next", "pixels"): torch.randint(0, 255, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                   "action": torch.randint(0, 6, (n, 1), device=dev, generator=g),
                   ("next", "reward"): torch.randn(n, 1, device=dev, generator=g),
                   ("next", "done"): torch.zeros(n, 1, dtype=torch.bool, device=dev),
                   ("next", "terminated"): torch.zeros(n, 1, dtype=torch.bool, device=dev)}, [n])
smp = PrioritizedSampler(1_000_000, 0.6, 0.4, device=dev)
smp.mark_update_range(0, 1_000_000, 1_000_000)
shape = (4096, 128, 1)
v, nv, r, lp, lm = (torch.randn(*shape, device=dev, generator=g) for _ in range(5))
term = torch.rand(*shape, device=dev, generator=g) < 0.02
done = term | (torch.rand(*shape, device=dev, generator=g) < 0.02)
L = 10_000_000
end = torch.rand(L, device=dev, generator=g) < 1 / 200
ids = torch.cumsum(end, 0)
table = torch.empty((3, L), dtype=torch.int64, device=dev)
counts = torch.zeros(2, dtype=torch.int64, device=dev)
ws = be.traj_workspace(L, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for it in range(3):
    flush.fill_(it)
    rb.extend(data)                                            # extend_kernel: 1024 Atari transitions + range update
    smp.mark_update_range(12345 + it, 4096, 1_000_000)         # tree_range_kernel
    flush.fill_(it + 1)
    vec_td_lambda_return_estimate(0.99, 0.95, nv, r, done, term)
    vtrace_advantage_estimate(0.99, lp, lm, v, nv, r, done, term)
    flush.fill_(it + 2)
    be.traj_table(end, False, L, True, -1, 64, True, table, counts, ws)
    be.traj_table(ids, True, L, True, -1, 64, True, table, counts, ws)
    k = int(counts[1])
    traj = torch.randint(k, (256,), device=dev, generator=g)
    be.slice_index(table[0], table[2], k, traj, torch.rand(256, device=dev, generator=g), 64, L)
torch.cuda.synchronize()
print("done")
