"""Write path: rb.extend through the fused rlb_extend launch vs the general path (slice copies / scatter
+ sorted-merge tree update), Atari-shaped transitions into a 300k-slot buffer; and the range update kernel alone vs the
general update kernel on the same slots.

    python profiles/exp_extend.py [capacity]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rl_b200.data import LazyTensorStorage, TensorDict, TensorDictPrioritizedReplayBuffer  # noqa: E402
from rl_b200.data.writers import RoundRobinWriter  # noqa: E402

dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000


def batch(n, g):
    return TensorDict({"pixels": torch.randint(0, 255, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                       ("next", "pixels"): torch.randint(0, 255, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                       "action": torch.randint(0, 6, (n, 1), device=dev, generator=g),
                       ("next", "reward"): torch.randn(n, 1, device=dev, generator=g),
                       ("next", "done"): torch.zeros(n, 1, dtype=torch.bool, device=dev),
                       ("next", "terminated"): torch.zeros(n, 1, dtype=torch.bool, device=dev)}, [n])


def timed(fn, iters):
    for _ in range(5):
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
    g = torch.Generator(device=dev).manual_seed(0)
    print(f"capacity {N}, device {torch.cuda.get_device_name(0)}")
    for n in (256, 1024, 4096):
        data = batch(n, g)
        row = sum(v.element_size() * v[0].numel() for v in data.values(True, True))
        res = {}
        for name in ("fused", "general"):
            rb = TensorDictPrioritizedReplayBuffer(alpha=0.6, beta=0.4, storage=LazyTensorStorage(N, device=dev),
                                                   batch_size=256)
            if name == "general":
                rb._writer._extend_fused = lambda *a: False
            rb.extend(data)
            res[name] = timed(lambda: rb.extend(data), 200 if n <= 1024 else 60)
            # device time: 10 consecutive extends (10 different slot ranges) captured in one CUDA graph
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(10):
                    rb.extend(data)
            res[name + "_graph"] = timed(gr.replay, 20) / 10
            del rb, gr
        bytes_ = 2 * n * row
        print(f"extend n={n:5d} row={row} B  eager: fused {res['fused']:7.1f} us, general {res['general']:7.1f} us | "
              f"graph-replayed: fused {res['fused_graph']:7.2f} us = {bytes_ / res['fused_graph'] / 1e3:6.0f} GB/s = "
              f"{n / res['fused_graph']:6.1f} M transitions/s, general {res['general_graph']:7.2f} us = "
              f"{bytes_ / res['general_graph'] / 1e3:6.0f} GB/s  ({res['general_graph'] / res['fused_graph']:.2f}x)")
    # tree part alone, graph-replayed so that host launch latency does not hide the kernels
    from rl_b200.data import PrioritizedSampler

    for n in (256, 1024, 4096, 65536):
        smp = PrioritizedSampler(1_000_000, 0.6, 0.4, device=dev)
        smp.mark_update_range(0, 1_000_000, 1_000_000)
        idx = (12345 + torch.arange(n, device=dev)) % 1_000_000
        out = {}
        for name, fn in (("range", lambda: smp.mark_update_range(12345, n, 1_000_000)),
                         ("general", lambda: smp.mark_update(idx))):
            if name == "general" and n > 1024:
                out[name] = timed(fn, 50)        # epoch-stamped path is not capturable
                continue
            fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(20):
                    fn()
            out[name] = timed(gr.replay, 20) / 20
        print(f"mark_update n={n:6d}: range kernel {out['range']:7.2f} us   general update {out['general']:7.2f} us")


if __name__ == "__main__":
    main()
