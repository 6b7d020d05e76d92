"""CUDA-graph capture of hot-path steps.

``rb.sample()`` / ``rb.update_priority()`` / ``GAE`` are a handful of short launches (8 us, 3 us, 4 us of device
time at the benchmark shapes): driven from Python they are launch-bound.  All of them are capture-safe -- no
host synchronisation, no data-dependent host control flow, allocations only through torch's caching allocator
-- so a training step built from them can be captured once and replayed with a single ``cudaGraphLaunch``.
This is the B200-idiomatic replacement for the reference's ``compilable=True`` / ``torch.compile`` route.

    step = CudaGraphStep(lambda: (rb.sample(), ...), generators=[rb._rng])
    out = step()          # replays; `out` are the SAME tensors every call (overwritten by the next replay)

Caveats (all checked or documented): the storage length, batch size, beta and every tensor passed in by
reference are frozen at capture time -- re-capture when they change; random draws are fresh on every replay
(the generator's Philox offset is advanced by the graph); ``update_priority`` batches up to 8192 items are one
replay-safe cluster launch, larger ones are applied under capture as consecutive chunks of 8192 (the epoch-stamped
path of very large eager batches is not replay-safe).
"""
from __future__ import annotations

from typing import Any, Callable, Sequence

import torch


class CudaGraphStep:
    """Capture ``fn()`` (no arguments; close over static inputs) into a CUDA graph and replay it on call."""

    def __init__(self, fn: Callable[[], Any], *, generators: Sequence[torch.Generator] = (), warmup: int = 3,
                 device=None, pool=None):
        if not torch.cuda.is_available():
            raise RuntimeError("CudaGraphStep needs a CUDA device")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.graph = torch.cuda.CUDAGraph()
        for g in generators:
            if g is not None:
                self.graph.register_generator_state(g)
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):  # lazy initialisation (tree allocation, func attributes) happens eagerly
                fn()
        self.stream.synchronize()
        with torch.cuda.graph(self.graph, stream=self.stream, pool=pool):
            self.outputs = fn()
        cur.wait_stream(self.stream)

    def pool(self):
        return self.graph.pool()

    def __call__(self):
        self.graph.replay()
        return self.outputs
