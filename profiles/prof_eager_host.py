"""Where the host time of the eager public calls goes: cProfile over rb.sample() / rb.update_priority() / rb.extend()."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
rb, g = bench.build_buffer(dev, 200_000, seed=0)
td_err = torch.rand(bench.BATCH, device=dev, generator=g)
for _ in range(50):
    b = rb.sample()
torch.cuda.synchronize()
for name, fn in (("sample", lambda: rb.sample()), ("update_priority", lambda: rb.update_priority(b.get("index"), td_err))):
    t0 = time.perf_counter()
    for _ in range(2000):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 2000 * 1e6:.1f} us per call (wall, eager)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(2000):
        fn()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(int(sys.argv[1]) if len(sys.argv) > 1 else 22)
