"""Probe: torch symmetric memory (NVLink peer buffers) availability on this box."""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty((1 << 20,), dtype=torch.uint8, device=dev)
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "rendezvous ok", type(hdl).__name__, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
t.fill_(rank + 1)
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.uint8)
print(rank, "peer value", int(peer[0].item()), flush=True)
# capture barrier in a graph
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(dev)
with torch.cuda.stream(s):
    hdl.barrier(channel=1)
    s.synchronize()
    try:
        with torch.cuda.graph(g, stream=s):
            hdl.barrier(channel=1)
        for _ in range(3):
            g.replay()
        s.synchronize()
        print(rank, "barrier captured+replayed ok", flush=True)
    except Exception as e:
        print(rank, "graph capture of barrier failed:", repr(e)[:200], flush=True)
dist.barrier(); dist.destroy_process_group()
