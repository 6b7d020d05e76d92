"""SliceSampler: trajectory-table build and per-sample cost on a 1M-slot ring of Atari-shaped steps.

    python profiles/exp_slice.py
"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rl_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def main():
    be = ops.backend()
    g = torch.Generator(device=dev).manual_seed(0)
    for L in (1_000_000, 10_000_000, 50_000_000):
        end = torch.rand(L, device=dev, generator=g) < 1 / 200
        ids = torch.cumsum(end, 0)
        table = torch.empty((3, L), dtype=torch.int64, device=dev)
        counts = torch.zeros(2, dtype=torch.int64, device=dev)
        ws = be.traj_workspace(L, dev)
        for name, sig, by_id in (("end flags", end, False), ("traj ids", ids, True)):
            for keep in (False, True):
                t = timed(lambda: be.traj_table(sig, by_id, L, True, -1, 64, keep, table, counts, ws), 20)
                passes = 3 if keep else 2
                n_all, n_long = counts.tolist()
                # bytes the passes move: flags 1 B/slot/pass; ids 8 B/slot once + the flag bytes written once and read
                # by the later passes; 24 B per table entry written
                moved = (passes * L if not by_id else 8 * L + L + (passes - 1) * L) + 24 * (n_long if keep else n_all)
                print(f"L={L:9d} {name:9s} filter={int(keep)}: table build {t:8.1f} us  ({passes} passes, "
                      f"{moved / 1e6:7.1f} MB moved = {moved / t / 1e3:7.1f} GB/s; {n_all} trajectories, {n_long} long enough)")

        # what the reference does for the same table (nonzero + roll by mask + boolean filter), torch ops on the GPU
        def ref_like():
            stop = end.nonzero().squeeze(-1)
            start = (stop.roll(1) + 1) % L
            length = stop - start + 1
            length = torch.where(length <= 0, length + L, length)
            keep = length >= 64
            return start[keep], stop[keep], length[keep]

        print(f"L={L:9d} torch-op restatement of the reference's table (syncs inside): {timed(ref_like, 10):8.1f} us")
        n_long = int(counts[1])
        S, T = 256, 64
        traj = torch.randint(n_long, (S,), device=dev, generator=g)
        u = torch.rand(S, device=dev, generator=g)
        t = timed(lambda: be.slice_index(table[0], table[2], n_long, traj, u, T, L), 200)
        print(f"L={L:9d} slice_index S={S} T={T}: {t:6.2f} us eager (one launch + 3 torch.empty)")
        del table, end, ids


def buffers():
    """rb.sample() through the public API: SliceSampler and PrioritizedSliceSampler on a 1M-slot ring, 32 slices x 64 steps
    of Atari-shaped frames (28 KB rows), trajectory table cached between writes."""
    from rl_b200.data import (LazyTensorStorage, PrioritizedSliceSampler, SliceSampler, TensorDict,
                              TensorDictReplayBuffer)

    L, S, T = 300_000, 32, 64
    g = torch.Generator(device=dev).manual_seed(1)
    for name, smp in (("SliceSampler", SliceSampler(num_slices=S, end_key=("next", "done"), cache_values=True)),
                      ("PrioritizedSliceSampler", PrioritizedSliceSampler(L, 0.6, 0.4, num_slices=S,
                                                                          end_key=("next", "done"), cache_values=True))):
        rb = TensorDictReplayBuffer(storage=LazyTensorStorage(L, device=dev), sampler=smp, batch_size=S * T, generator=g)
        for lo in range(0, L, 50_000):
            n = 50_000
            rb.extend(TensorDict({"pixels": torch.randint(0, 255, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                                  ("next", "done"): torch.rand(n, 1, device=dev, generator=g) < 0.005}, [n]))
        rb.sample()
        t = timed(rb.sample, 100)
        bytes_ = 2 * S * T * 28225
        print(f"{name:24s} rb.sample() {S}x{T} steps of 28 KB: {t:7.1f} us eager ({bytes_ / t / 1e3:6.0f} GB/s of rows, "
              f"{S * T / t:5.1f} M steps/s)")
        del rb


if __name__ == "__main__":
    main()
    buffers()
