#!/usr/bin/env python
"""bench.py -- the replay-and-advantage hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input (BASELINE.json configs[1] + [2]):

    1. batch = rb.sample(256)            PER sample on a 1M-capacity buffer of Atari transitions (pixels and
                                         next pixels u8[4,84,84], action i64, reward f32, done, terminated):
                                         torch.rand -> rlb_per_sample -> rlb_gather of all six leaves
    2. rb.update_priority(index, |td|)   (p+eps)^alpha, last-writer-wins scatter, touched-ancestor recompute
    3. adv, tgt = GAE([4096, 128])       gamma=.99, lmbda=.95, done/terminated ~ Bernoulli(.02)

`value` = (256 sampled + 4096*128 GAE) transitions per step / device time per step, whole job.  With N > 1
the buffer is sharded by capacity (one 1M shard per rank), every rank draws 256 from its shard
and ONE NCCL all-gather assembles the 256*N global minibatch on every rank; GAE rows are split by rank
(weak scaling: per-GPU work is fixed).

Timing: CUDA events on the launching stream around exactly K steps after W warm-ups, barrier +
synchronize on both sides, max over ranks.  Inputs are larger than L2 (the storage is 56 GB; the GAE inputs
rotate through a ring of sets larger than the 126 MB L2).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

ALPHA, BETA, GAMMA, LMBDA = 0.6, 0.4, 0.99, 0.95
CAPACITY, BATCH = 1_000_000, 256
GAE_ROWS, GAE_T = 4096, 128
ROW_BYTES = 2 * 4 * 84 * 84 + 8 + 4 + 1 + 1  # 56 462 B per Atari transition
TRANSITIONS_PER_STEP = BATCH + GAE_ROWS * GAE_T
METRIC = "transitions/sec sampled+GAE at 1M buffer / Atari frames"


# ------------------------------------------------------------------------------------------- helpers
def peaks() -> tuple[float, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def timed(fn, steps: int, warmup: int, sync_all) -> float:
    """ms per step over exactly `steps` calls of fn(i), CUDA events on the current stream."""
    for i in range(warmup):
        fn(i)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    sync_all()
    return e0.elapsed_time(e1) / steps


# ------------------------------------------------------------------------------------------- our arm
def build_buffer(dev, capacity: int, seed: int):
    """1M-capacity HBM-resident PER buffer of Atari-shaped transitions, filled with synthetic data."""
    from rl_b200.data import LazyTensorStorage, TensorDict, TensorDictPrioritizedReplayBuffer

    g = torch.Generator(device=dev).manual_seed(seed)
    rb = TensorDictPrioritizedReplayBuffer(alpha=ALPHA, beta=BETA, storage=LazyTensorStorage(capacity, device=dev),
                                           batch_size=BATCH, generator=g)
    chunk = 50_000
    for lo in range(0, capacity, chunk):
        n = min(chunk, capacity - lo)
        td = TensorDict({
            "pixels": torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
            "action": torch.randint(0, 18, (n, 1), device=dev, generator=g),
            "next": {"pixels": torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                     "reward": torch.randn(n, device=dev, generator=g),
                     "done": torch.rand(n, 1, device=dev, generator=g) < 0.01,
                     "terminated": torch.rand(n, 1, device=dev, generator=g) < 0.01},
            "td_error": torch.rand(n, device=dev, generator=g),
        }, [n])
        rb.extend(td)
    return rb, g


def gae_ring(dev, rows: int, T: int, g, min_bytes: int = 160 << 20):
    """Input sets for GAE whose total footprint exceeds L2 so that every call reads cold data."""
    per = rows * T * 14
    n = max(2, -(-min_bytes // per))
    n += n % 2  # even: the sharded buffer alternates its two receive buffers from one step to the next
    ring = []
    for _ in range(n):
        v, nv, r = (torch.randn(rows, T, 1, device=dev, generator=g) for _ in range(3))
        term = torch.rand(rows, T, 1, device=dev, generator=g) < 0.02
        done = term | (torch.rand(rows, T, 1, device=dev, generator=g) < 0.02)
        ring.append((v, nv, r, done, term))
    return ring


def kernel_roofline(dev, rb, ring, hbm_peak: float, peak_src: str) -> dict:
    """Average launch duration of the dominant kernel (the gather: 28.9 MB/launch vs 11.5 MB for GAE), timed
    with CUDA events around a CUDA-graph replay of back-to-back launches on cold rows (no host in the loop)."""
    from rl_b200 import ops

    be = ops.backend()
    st = rb.storage
    n_launch = 20
    g = torch.Generator(device=dev).manual_seed(123)
    idxs = [torch.randint(0, len(st), (BATCH,), device=dev, generator=g) for _ in range(n_launch)]
    stream = torch.cuda.Stream(dev)
    out = {}
    with torch.cuda.stream(stream):
        # warm-up outside capture
        for ix in idxs[:3]:
            be.gather(st._leaves, ix, len(st))
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            keep = [be.gather(st._leaves, ix, len(st)) for ix in idxs]
        ms = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            graph.replay()
            e1.record(stream)
            stream.synchronize()
            ms.append(e0.elapsed_time(e1) / n_launch)
        del keep
        us = sorted(ms)[len(ms) // 2] * 1e3
        row_bytes = sum(l[0].numel() * l.element_size() for l in st._leaves)  # every stored leaf, per transition
        alg = BATCH * row_bytes * 2 + BATCH * 8
        traffic = None
        tf = ROOT / "profiles" / "traffic_r1.json"
        if tf.exists():  # from the committed `ncu --set full` capture (profiles/ncu_full_r1.csv), per launch
            tj = json.loads(tf.read_text())["gather_kernel_B256"]
            traffic = tj["dram_read_bytes"] + tj["dram_write_bytes"]
        out = {"bound": "hbm", "kernel": f"gather_kernel<gather> (rlb_gather, {len(st._leaves)} leaves, B=256)",
               "row_bytes": row_bytes,
               "achieved": round(alg / us / 1e3, 1), "peak": hbm_peak, "unit": "GB/s",
               "frac": round(alg / us / 1e3 / hbm_peak, 4), "traffic": traffic, "peak_source": peak_src,
               "algorithmic_bytes": alg, "us_per_launch": round(us, 3)}
        # the same kernel over a batch sweep (graph-replayed launches on cold rows): fixed cost + streaming rate
        sweep = {}
        for bb, reps in ((1024, 8), (4096, 4), (16384, 2)):
            idx_b = [torch.randint(0, len(st), (bb,), device=dev, generator=g) for _ in range(reps)]
            be.gather(st._leaves, idx_b[0], len(st))
            stream.synchronize()
            gb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gb, stream=stream):
                keep = [be.gather(st._leaves, ix, len(st)) for ix in idx_b]
            gb.replay()
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            gb.replay()
            e1.record(stream)
            stream.synchronize()
            del keep, gb
            us_b = e0.elapsed_time(e1) * 1e3 / reps
            alg_b = bb * row_bytes * 2 + bb * 8
            sweep[f"B{bb}"] = {"achieved": round(alg_b / us_b / 1e3, 1), "frac": round(alg_b / us_b / 1e3 / hbm_peak, 4),
                               "us_per_launch": round(us_b, 1)}
        out["batch_sweep"] = sweep
        out["steady_state_B16384"] = sweep["B16384"]
        # GAE kernel, same method
        v, nv, r, d, t = ring[0]
        graph2 = torch.cuda.CUDAGraph()
        d8 = [(x[3].view(torch.uint8), x[4].view(torch.uint8)) for x in ring]
        for i in range(2):
            be.gae(ring[i][0], ring[i][1], ring[i][2], d8[i][0], d8[i][1], GAMMA, GAMMA * LMBDA, GAE_ROWS, GAE_T, 1)
        stream.synchronize()
        with torch.cuda.graph(graph2, stream=stream):
            keep = [be.gae(x[0], x[1], x[2], dd[0], dd[1], GAMMA, GAMMA * LMBDA, GAE_ROWS, GAE_T, 1)
                    for x, dd in zip(ring, d8)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        graph2.replay()
        stream.synchronize()
        e0.record(stream)
        graph2.replay()
        e1.record(stream)
        stream.synchronize()
        us_g = e0.elapsed_time(e1) * 1e3 / len(ring)
        alg_g = GAE_ROWS * GAE_T * 22
        out["gae_kernel"] = {"achieved": round(alg_g / us_g / 1e3, 1), "frac": round(alg_g / us_g / 1e3 / hbm_peak, 4),
                             "us_per_launch": round(us_g, 3), "algorithmic_bytes": alg_g}
        del keep
    return out


def write_path_probe(dev, rb, g, hbm_peak: float) -> dict | None:
    """SURVEY 8(f)-1 beside the headline: rb.extend of 1024 Atari transitions (rows + default priorities, one rlb_extend
    launch), ten consecutive writer batches captured in one CUDA graph and replayed.  Never fatal to the bench."""
    try:
        from rl_b200.data import TensorDict

        n = 1024
        td = TensorDict({
            "pixels": torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
            "action": torch.randint(0, 18, (n, 1), device=dev, generator=g),
            "next": {"pixels": torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                     "reward": torch.randn(n, device=dev, generator=g),
                     "done": torch.zeros(n, 1, dtype=torch.bool, device=dev),
                     "terminated": torch.zeros(n, 1, dtype=torch.bool, device=dev)},
            "td_error": torch.rand(n, device=dev, generator=g),
        }, [n])
        row = sum(l[0].numel() * l.element_size() for l in rb.storage._leaves if l is not None)
        rb.extend(td, update_priority=False)
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(dev)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(gr, stream=side):
                for _ in range(10):
                    rb.extend(td, update_priority=False)
        torch.cuda.synchronize(dev)
        for _ in range(3):
            gr.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        a.record()
        for _ in range(20):
            gr.replay()
        b.record()
        torch.cuda.synchronize(dev)
        us = a.elapsed_time(b) * 1e3 / 200
        gbs = 2 * n * row / us / 1e3
        return {"extend_n": n, "extend_us": round(us, 2), "extend_GBps": round(gbs, 1), "extend_frac_of_hbm": round(gbs / hbm_peak, 3),
                "extend_transitions_per_s": round(n / (us * 1e-6), 1), "row_bytes": row}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:200]}


def run_ours(args) -> dict:
    from rl_b200 import ops
    from rl_b200.graphs import CudaGraphStep

    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world > 1
    if distributed:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    be = ops.backend()
    gbatch = BATCH * world
    if distributed:
        rb, g = build_sharded(dev, CAPACITY, world, rank)
    else:
        rb, g = build_buffer(dev, CAPACITY, seed=rank)
    ring = gae_ring(dev, GAE_ROWS, GAE_T, g)
    ring8 = [(v, nv, r, d.view(torch.uint8), t.view(torch.uint8)) for v, nv, r, d, t in ring]
    R = len(ring)
    td_err = torch.rand(gbatch, device=dev, generator=g)
    gs = (float(torch.tensor(GAMMA)), float(torch.tensor(GAMMA) * torch.tensor(LMBDA)))

    def sync_all():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    side_gae, side_upd = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    smp = rb.sampler
    smp.record_index_event = True
    # optional: each sample also draws the NEXT sample's uniforms (same torch.rand calls, same order), taking the RNG
    # kernel off the chain rand -> per_sample -> update.  Measured: 37.6 -> 37.1 us per step, so the headline keeps the
    # plain call sequence
    smp.predraw = os.environ.get("RLB_BENCH_PREDRAW", "0") != "0"

    def make_step(slot: int):
        v, nv, r, d8, t8 = ring8[slot]

        def step():
            # three independent chains of the same step run on three streams (fork / join with events):
            #   main      rand -> per_sample -> gather [-> all-gather]
            #   side_upd  update_priority(index): needs the sampled indices only, overlaps the gather
            #   side_gae  GAE of this step's rollout: independent of the replay path
            main = torch.cuda.current_stream(dev)
            side_gae.wait_stream(main)
            with torch.cuda.stream(side_gae):
                a, tg = be.gae(v, nv, r, d8, t8, gs[0], gs[1], GAE_ROWS, GAE_T, 1)
            batch = rb.sample()
            index = batch.get("index")
            if distributed:
                side_upd.wait_stream(main)           # global indices exist only after the all-gather
            else:
                side_upd.wait_event(smp.index_ready)
            with torch.cuda.stream(side_upd):
                rb.update_priority(index, td_err)    # fused pow + tree write-back
            main.wait_stream(side_upd)
            main.wait_stream(side_gae)
            return batch, a, tg

        return step

    def make_eager_step(slot: int):
        v, nv, r, d8, t8 = ring8[slot]

        def step():  # the plain single-stream call sequence a Python training loop would issue
            batch = rb.sample()
            rb.update_priority(batch.get("index"), td_err)
            a, tg = be.gae(v, nv, r, d8, t8, gs[0], gs[1], GAE_ROWS, GAE_T, 1)
            return batch, a, tg

        return step

    steps = [make_step(i) for i in range(R)]
    eager_steps = [make_eager_step(i) for i in range(R)]

    # ---- (1) eager: every call goes through the Python API
    clocks = ClockSampler(local) if rank == 0 else None
    ms_eager = timed(lambda i: eager_steps[i % R](), args.steps, args.warmup, sync_all)

    # ---- (2) the same steps captured once into CUDA graphs (one per GAE input set) and replayed.  With N > 1 the
    # NCCL all-gather is issued eagerly between two captured halves: [local draw into the send buffer] and
    # [weights + priority write-back + GAE].
    graphs, graph_err, ms_graph = None, None, None
    try:
        gen = rb.sampler._rng
        if not distributed:
            graphs = [CudaGraphStep(st, generators=[gen], warmup=1) for st in steps]
            ms_graph = timed(lambda i: graphs[i % R](), args.steps, args.warmup, sync_all)
        elif rb._use_nvlink(dev):
            # sharded + NVLink transport: the gather kernel itself broadcasts the rows into every rank's receive
            # buffer and a signal-pad barrier closes the exchange -- no NCCL call, so the WHOLE step is one graph.
            # Receive buffers are double-buffered: slot i uses buffer i % 2 (R is even), fixed at capture time.
            def make_dist_step(slot: int):
                v, nv, r, d8, t8 = ring8[slot]

                def step():
                    main = torch.cuda.current_stream(dev)
                    side_gae.wait_stream(main)
                    with torch.cuda.stream(side_gae):
                        a, tg = be.gae(v, nv, r, d8, t8, gs[0], gs[1], GAE_ROWS, GAE_T, 1)
                    rb.local_draw()
                    rb.exchange()
                    batch = rb.finalize()
                    rb.update_priority(batch.get("index"), td_err)
                    main.wait_stream(side_gae)
                    return batch, a, tg

                return step

            graphs = [CudaGraphStep(make_dist_step(i), generators=[gen], warmup=2) for i in range(R)]  # 3 calls/slot
            ms_graph = timed(lambda i: graphs[i % R](), args.steps, args.warmup, sync_all)
        else:
            draw = CudaGraphStep(lambda: rb.local_draw(static_buffers=True), generators=[gen], warmup=1)

            def make_tail(slot: int):
                v, nv, r, d8, t8 = ring8[slot]

                def tail():
                    batch = rb.finalize()
                    rb.update_priority(batch.get("index"), td_err)
                    return batch, be.gae(v, nv, r, d8, t8, gs[0], gs[1], GAE_ROWS, GAE_T, 1)

                return tail

            rb.exchange()
            tails = [CudaGraphStep(make_tail(i), warmup=1) for i in range(R)]

            def graph_step(i):
                draw()
                rb.exchange()
                return tails[i % R]()

            ms_graph = timed(graph_step, args.steps, args.warmup, sync_all)
    except Exception as err:
        import traceback

        graph_err = f"{type(err).__name__}: {err}"[:300]
        print(f"[bench rank {rank}] CUDA-graph path failed, falling back to eager: {graph_err}\n"
              + "".join(traceback.format_exc().splitlines(True)[-6:]), file=sys.stderr, flush=True)
        graphs, ms_graph = None, None
        # a capture that died half-way leaves the generator in capture mode: give the sampler a fresh one
        torch.cuda.synchronize()
        fresh = torch.Generator(device=dev).manual_seed(4242 + rank)
        if hasattr(rb, "local"):
            rb.local.set_rng(fresh)
        else:
            rb.set_rng(fresh)
        g = fresh
    clk = clocks.stop() if clocks else None

    # ---- sub-metrics (eager, same method)
    ms_sample = timed(lambda i: rb.sample(), args.steps, 3, sync_all)
    idx_pool = [torch.randint(0, CAPACITY * world, (gbatch,), device=dev, generator=g) for _ in range(8)]
    ms_update = timed(lambda i: rb.update_priority(idx_pool[i % 8], td_err), args.steps, 3, sync_all)

    def gae_only(i):
        v, nv, r, d8, t8 = ring8[i % R]
        be.gae(v, nv, r, d8, t8, gs[0], gs[1], GAE_ROWS, GAE_T, 1)

    ms_gae = timed(gae_only, args.steps, 3, sync_all)

    # ---- (3) end to end with HOST buffers: H2D of the step's inputs and D2H of the step's results every step.
    # Three independent lanes (stream + pinned buffers + device staging) rotate so that one step's D2H overlaps
    # the next steps' H2D and compute (PCIe is full duplex); every step still synchronises on its own results.
    pin = lambda t: t.cpu().pin_memory()
    sample0 = rb.sample()
    out_keys = list(sample0.keys(True, True))
    # N > 1: every rank delivers ITS rows of the gathered batch to the host (the job as a whole lands the global
    # minibatch in host memory once), plus its own GAE outputs
    own = slice(rank * BATCH, (rank + 1) * BATCH) if distributed else slice(None)
    lanes = []
    # N > 1: one lane -- the sharded buffer's two receive buffers are overwritten by the PEERS' next-but-one draw,
    # which is only ordered against work on the sampling stream
    n_lanes = 1 if distributed else 3
    for lane in range(n_lanes):
        s = torch.cuda.Stream(dev)
        host_in = tuple(pin(x) for x in ring[lane % R])
        lanes.append({
            "stream": s, "host_in": host_in, "host_td": pin(td_err),
            "dev_in": [torch.empty_like(x) for x in ring[lane % R]], "dev_td": torch.empty_like(td_err),
            "host_out": {k: torch.empty(sample0.get(k)[own].shape, dtype=sample0.get(k).dtype).pin_memory()
                         for k in out_keys},
            "host_adv": torch.empty(GAE_ROWS, GAE_T, 1).pin_memory(), "host_tgt": torch.empty(GAE_ROWS, GAE_T, 1).pin_memory(),
            "done": torch.cuda.Event(),
        })
    h2d = sum(x.numel() * x.element_size() for x in lanes[0]["host_in"]) + td_err.numel() * 4
    d2h = sum(v.numel() * v.element_size() for v in lanes[0]["host_out"].values()) + 2 * GAE_ROWS * GAE_T * 4

    def e2e_body(L):
        for dst, src in zip(L["dev_in"], L["host_in"]):
            dst.copy_(src, non_blocking=True)
        L["dev_td"].copy_(L["host_td"], non_blocking=True)
        batch = rb.sample()
        rb.update_priority(batch.get("index"), L["dev_td"])
        di = L["dev_in"]
        a, tg = be.gae(di[0], di[1], di[2], di[3].view(torch.uint8), di[4].view(torch.uint8), gs[0], gs[1],
                       GAE_ROWS, GAE_T, 1)
        for k, hv in L["host_out"].items():
            hv.copy_(batch.get(k)[own], non_blocking=True)
        L["host_adv"].copy_(a, non_blocking=True)
        L["host_tgt"].copy_(tg, non_blocking=True)

    def e2e_step(i: int):
        L = lanes[i % n_lanes]
        L["done"].synchronize()            # the caller consumed this lane's previous results
        with torch.cuda.stream(L["stream"]):
            e2e_body(L)
            L["done"].record()

    def timed_e2e(steps_, warmup_):
        for i in range(warmup_):
            e2e_step(i)
        sync_all()
        cur = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for L in lanes:
            L["stream"].wait_event(e0)
        for i in range(steps_):
            e2e_step(warmup_ + i)
        for L in lanes:
            cur.wait_stream(L["stream"])
        e1.record(cur)
        sync_all()
        return e0.elapsed_time(e1) / steps_

    ms_e2e = timed_e2e(args.steps, args.warmup)

    # ---- max over ranks
    vals = [ms_eager, ms_graph if ms_graph is not None else -1.0, ms_e2e, ms_sample, ms_update, ms_gae]
    if distributed:
        t = torch.tensor(vals, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = t.tolist()
    ms_eager, ms_graph, ms_e2e, ms_sample, ms_update, ms_gae = vals
    if ms_graph is not None and ms_graph <= 0:
        ms_graph = None

    hbm_peak, peak_src = peaks()
    transport = None
    if distributed:
        transport = ("gather kernel broadcasts over NVLink peer memory + signal-pad barrier" if rb._symm not in (None, False)
                     else "NCCL all-gather issued eagerly between two graphs")
    result = None
    if rank == 0:
        roof = kernel_roofline(dev, rb, ring, hbm_peak, peak_src) if world == 1 else None
        cpu = cpu_baseline_run(steps=None) if world == 1 else None
        per_step = TRANSITIONS_PER_STEP * world
        ms = ms_graph if ms_graph is not None else ms_eager
        n_leaves = len(rb.storage._leaves)
        result = {
            "metric": METRIC, "value": round(per_step / (ms * 1e-3), 1), "unit": "transitions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C2 PER sample+update B=256 @1M Atari transitions ({n_leaves} leaves) + C3 GAE [4096,128]",
                       "capacity_per_gpu": CAPACITY, "batch_per_gpu": BATCH, "gae_shape": [GAE_ROWS, GAE_T, 1],
                       "arithmetic": "fp32 priorities / trees / GAE, byte-exact u8 row moves, int64 indices",
                       "alpha": ALPHA, "beta": BETA, "gamma": GAMMA, "lmbda": LMBDA,
                       "launch": ("cuda_graph replay of the public-API step" + (f" ({transport})" if world > 1 else "")) if ms_graph is not None else "eager python API",
                       "parallelism": f"capacity-sharded x{world}, {transport}" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (56 GB storage, random rows; GAE inputs rotate through >160 MB)"},
            "e2e": {"value": round(per_step / (ms_e2e * 1e-3), 1), "unit": "transitions/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e, 5),
                    "note": f"eager python API, pinned host buffers, {n_lanes} lane(s): with several lanes D2H overlaps the next steps' H2D + compute"},
            "gpu_launches": (4 if world == 1 else 6) * args.steps,  # ours per step: per_sample, gather, update, gae [+ pack, weights]
            "breakdown": {"eager_ms_per_step": round(ms_eager, 5),
                          "graph_ms_per_step": None if ms_graph is None else round(ms_graph, 5),
                          "graph_error": graph_err,
                          "eager_value": round(per_step / (ms_eager * 1e-3), 1),
                          "sample_us": round(ms_sample * 1e3, 2), "update_priority_us": round(ms_update * 1e3, 2),
                          "gae_us": round(ms_gae * 1e3, 2),
                          "sample_transitions_per_s": round(gbatch / (ms_sample * 1e-3), 1),
                          "gae_transitions_per_s": round(GAE_ROWS * GAE_T * world / (ms_gae * 1e-3), 1)},
            "clocks": clk,
        }
        if roof:
            result["roofline"] = roof
        if cpu:
            result["cpu_baseline"] = cpu
        if world == 1:
            result["breakdown"]["write_path"] = write_path_probe(dev, rb, g, hbm_peak)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return result


def build_sharded(dev, capacity_per_rank: int, world: int, rank: int):
    """One 1M shard per rank of a capacity-sharded buffer (weak scaling), filled with synthetic transitions."""
    from rl_b200.data import TensorDict
    from rl_b200.data.sharded import ShardedPrioritizedReplayBuffer

    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    rb = ShardedPrioritizedReplayBuffer(alpha=ALPHA, beta=BETA, capacity=capacity_per_rank * world,
                                        batch_size=BATCH * world, device=dev, generator=g)
    chunk = 50_000
    for lo in range(0, capacity_per_rank, chunk):
        n = min(chunk, capacity_per_rank - lo)
        rb.extend(TensorDict({
            "pixels": torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
            "action": torch.randint(0, 18, (n, 1), device=dev, generator=g),
            "next": {"pixels": torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
                     "reward": torch.randn(n, device=dev, generator=g),
                     "done": torch.rand(n, 1, device=dev, generator=g) < 0.01,
                     "terminated": torch.rand(n, 1, device=dev, generator=g) < 0.01},
            "td_error": torch.rand(n, device=dev, generator=g)}, [n]))
    return rb, g


# ------------------------------------------------------------------------------------------- reference arm
def cpu_baseline_run(steps: int | None, warmup: int = 2) -> dict:
    """The reference's own CPU path on the host cores: compiled reference segment trees (oracle/_ref/cpu) under
    the restated sampler glue, aten::index gather on CPU tensors, and the faster of the reference's two GAE code
    paths.  Bounded sample: a 20k-row CPU storage (gather cost is per row, not per capacity) and ~10-20 s of work."""
    from oracle import gae_torch
    from oracle import per_oracle as po
    from oracle.ref_loader import reference_trees

    cores = os.cpu_count() or 1
    factory = reference_trees("cpu")
    kind = "reference" if factory is not None else "port"
    smp = po.OraclePrioritizedSampler(CAPACITY, ALPHA, BETA, tree_factory=factory)
    g = torch.Generator().manual_seed(0)
    rows = 20_000
    smp.update_priority(torch.arange(CAPACITY), torch.rand(CAPACITY, generator=g))
    store = {"pixels": torch.randint(0, 256, (rows, 4, 84, 84), dtype=torch.uint8, generator=g),
             "next_pixels": torch.randint(0, 256, (rows, 4, 84, 84), dtype=torch.uint8, generator=g),
             "action": torch.randint(0, 18, (rows, 1), generator=g), "reward": torch.randn(rows, generator=g),
             "done": torch.rand(rows, 1, generator=g) < 0.01, "terminated": torch.rand(rows, 1, generator=g) < 0.01}
    v, nv, r = (torch.randn(GAE_ROWS, GAE_T, 1, generator=g) for _ in range(3))
    term = torch.rand(GAE_ROWS, GAE_T, 1, generator=g) < 0.02
    done = term | (torch.rand(GAE_ROWS, GAE_T, 1, generator=g) < 0.02)
    gm, lm = torch.tensor(GAMMA), torch.tensor(LMBDA)
    td = torch.rand(BATCH, generator=g)

    def t_of(fn, n):
        fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n

    # give the reference the thread count it runs fastest with (many-core hosts oversubscribe the small ops of
    # the GAE code paths; aten::index parallelises over rows): tuned per phase on the bounded sample
    cands = sorted({c for c in (1, 4, 8, 16, 32, 64, cores) if c <= cores})
    ridx0 = torch.randint(0, rows, (BATCH,), generator=g)
    best_gae, best_gather = (None, None, 1e9), (None, 1e9)
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            for name, fn in (("loop", gae_torch.loop_gae), ("vec", gae_torch.vec_gae)):
                t = t_of(lambda: fn(gm, lm, v, nv, r, done, term), 2)
                if t < best_gae[2]:
                    best_gae = (name, c, t)
            t = t_of(lambda: {k: x[ridx0] for k, x in store.items()}, 3)
            if t < best_gather[1]:
                best_gather = (c, t)
    gae_fn = gae_torch.loop_gae if best_gae[0] == "loop" else gae_torch.vec_gae
    t_loop = t_vec = best_gae[2]

    def step():
        torch.set_num_threads(best_gather[0])
        idx, w = smp.sample(CAPACITY, BATCH, generator=g)
        ridx = idx % rows  # bounded storage: same number of random 28 KB rows touched
        batch = {k: x[ridx] for k, x in store.items()}
        smp.update_priority(idx, td)
        torch.set_num_threads(best_gae[1])
        with torch.no_grad():
            a, t = gae_fn(gm, lm, v, nv, r, done, term)
        return batch, a, t

    for _ in range(warmup):
        step()
    if steps is None:
        t1 = t_of(step, 2)
        steps = max(5, min(400, int(12.0 / max(t1, 1e-4))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": round(TRANSITIONS_PER_STEP / dt, 1), "unit": "transitions/s",
            "cores": max(best_gather[0], best_gae[1]), "host_cores": cores, "kind": kind,
            "ms_per_step": round(dt * 1e3, 3), "steps": steps,
            "sample": (f"{steps} steps of [reference C++ SumSegmentTree/MinSegmentTree sample+update B=256 @1M "
                       f"(single-threaded by construction) + aten::index gather of 256 Atari transitions from a "
                       f"{rows}-row CPU storage ({best_gather[0]} threads, {best_gather[1] * 1e3:.2f} ms) + "
                       f"{best_gae[0]} GAE [4096,128] ({best_gae[1]} threads, {best_gae[2] * 1e3:.1f} ms)]; thread "
                       f"counts are the fastest of {cands} for each phase")}


def run_reference(args) -> dict | None:
    rank, world, _ = dist_env()
    if rank != 0:
        return None
    cpu = cpu_baseline_run(steps=args.steps, warmup=args.warmup)
    return {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "transitions/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "C2 PER sample+update B=256 @1M Atari transitions (56462 B/row) + C3 GAE [4096,128]",
                       "note": "reference CPU implementation on the host cores; rank 0 only"},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def _watchdog(seconds: float) -> None:
    """A hung collective must not hang the caller: after `seconds` the process prints what it knows and exits."""
    def fire():
        rank = int(os.environ.get("RANK", 0))
        if rank == 0:
            print(json.dumps({"metric": METRIC, "error": f"bench.py watchdog fired after {seconds:.0f} s",
                              "n_gpus": int(os.environ.get("WORLD_SIZE", 1))}), flush=True)
        os._exit(3)

    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    _watchdog(float(os.environ.get("RLB_BENCH_TIMEOUT", "840")))
    if args.impl == "reference":
        res = run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; rl_b200 has no CPU fallback (use --impl reference for the CPU arm)")
        res = run_ours(args)
    if res is not None:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
