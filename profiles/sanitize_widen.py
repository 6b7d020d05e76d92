"""Small invocations of every SURVEY 8(f) kernel for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool memcheck  python profiles/sanitize_widen.py
    compute-sanitizer --tool racecheck python profiles/sanitize_widen.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rl_b200 import ops  # noqa: E402
from rl_b200.data import (LazyTensorStorage, PrioritizedSampler, PrioritizedSliceSampler, SliceSampler,  # noqa: E402
                          TensorDict, TensorDictPrioritizedReplayBuffer, TensorDictReplayBuffer)
from rl_b200.objectives.value import vec_td_lambda_return_estimate, vtrace_advantage_estimate  # noqa: E402

dev = torch.device("cuda", 0)
be = ops.backend()
g = torch.Generator(device=dev).manual_seed(0)

# write path: fused extend with wrap-around, rows wide and narrow; range update alone
rb = TensorDictPrioritizedReplayBuffer(alpha=0.6, beta=0.4, storage=LazyTensorStorage(3000, device=dev), batch_size=64)
for n in (1000, 1500, 1200, 7):
    rb.extend(TensorDict({"pixels": torch.randint(0, 255, (n, 2, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                          "action": torch.randint(0, 6, (n, 1), device=dev, generator=g),
                          "reward": torch.randn(n, device=dev, generator=g)}, [n]))
    b = rb.sample()
    rb.update_priority(b.get("index"), torch.rand(64, device=dev, generator=g))
smp = PrioritizedSampler(100_000, 0.6, 0.4, device=dev)
for start, n in ((0, 100_000), (99_990, 50), (123, 1), (5000, 70_000)):
    smp.mark_update_range(start, n, 100_000)
smp._sum_tree.load_leaves(torch.rand(100_000, device=dev, generator=g))      # tiled rebuild

# scans
shape = (64, 200, 1)
v, nv, r, lp, lm = (torch.randn(*shape, device=dev, generator=g) for _ in range(5))
term = torch.rand(*shape, device=dev, generator=g) < 0.05
done = term | (torch.rand(*shape, device=dev, generator=g) < 0.05)
vec_td_lambda_return_estimate(0.99, 0.95, nv, r, done, term)
vtrace_advantage_estimate(0.99, lp, lm, v, nv, r, done, term)
vtrace_advantage_estimate(0.99, lp.squeeze(-1), lm.squeeze(-1), v.squeeze(-1), nv.squeeze(-1), r.squeeze(-1),
                          done.squeeze(-1), term.squeeze(-1), time_dim=-1)

# trajectory table (flags and ids, filter on/off, odd length), slice expansion, slice samplers
for L in (1, 37, 16_385, 100_003):
    end = torch.rand(L, device=dev, generator=g) < 0.03
    ids = torch.cumsum(end, 0)
    table = torch.empty((3, L), dtype=torch.int64, device=dev)
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    ws = be.traj_workspace(L, dev)
    for sig, by_id in ((end, False), (ids, True)):
        for keep in (False, True):
            be.traj_table(sig, by_id, L, True, L // 2, 8, keep, table, counts, ws)
            be.traj_table(sig, by_id, L, False, -1, 8, keep, table, counts, ws)
L = 20_000
for cls, kw in ((SliceSampler, dict(strict_length=False)), (SliceSampler, dict(strict_length=False, pad_output=True)),
                (SliceSampler, {}), (PrioritizedSliceSampler, {})):
    args = (L, 0.6, 0.4) if cls is PrioritizedSliceSampler else ()
    rb = TensorDictReplayBuffer(storage=LazyTensorStorage(L, device=dev), batch_size=16 * 24, generator=g,
                                sampler=cls(*args, num_slices=16, end_key=("next", "done"), **kw))
    rb.extend(TensorDict({"t": torch.arange(L, device=dev).reshape(L, 1),
                          ("next", "done"): torch.rand(L, 1, device=dev, generator=g) < 0.05}, [L]))
    for _ in range(2):
        rb.sample()
torch.cuda.synchronize()
print("sanitize_widen done")
