"""clock64() stamps at the phase boundaries of the single-launch update kernel (profiling aid)."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rl_b200 import ops
from rl_b200.data import PrioritizedSampler
dev = torch.device("cuda", 0)
be = ops.backend()
be.L.rlb_debug_set_tick_buffer.argtypes = [ctypes.c_void_p]
N = 1_000_000
for B in (256, 1024):
    smp = PrioritizedSampler(N, 0.6, 0.4, device=dev)
    smp.update_priority(torch.arange(N, device=dev), torch.rand(N, device=dev))
    ticks = torch.zeros(64, dtype=torch.int64, device=dev)
    be.L.rlb_debug_set_tick_buffer(ticks.data_ptr())
    for it in range(4):
        idx = torch.randint(0, N, (B,), device=dev)
        p = torch.rand(B, device=dev)
        torch.cuda.synchronize()
        smp.update_priority(idx, p)
        torch.cuda.synchronize()
        t = ticks.tolist()
        d = [t[i + 1] - t[i] for i in range(7)]
        print(f"UPD B={B} it={it}: start {d[0]}  keys {d[1]}  sort {d[2]}  heads+wait2 {d[3]}  init {d[4]}  climb {d[5]}  dense+store {d[6]}  "
              f"[deposit {t[12]-t[6]} arrive3 {t[13]-t[12]} dense {t[14]-t[13]} store {t[7]-t[14]}]  total {t[7]-t[0]} cycles | globaltimer ns: "
              f"leader start 0, leader tick7 {t[33]-t[32]}, leader end {t[34]-t[32]}; helper1 start {t[40]-t[32]}, stores issued {t[41]-t[32]}, "
              f"phase3 seen {t[42]-t[32]}, rows stored {t[43]-t[32]}")
        u = torch.rand(B, device=dev)
        torch.cuda.synchronize()
        be.per_sample(smp._sum_tree.values, smp._min_tree.values, N, smp._sum_tree.capacity, N, u, 0.4, True)
        torch.cuda.synchronize()
        t = ticks.tolist()
        print(f"SMP B={B} it={it}: psum+sync {t[9]-t[8]}  descent {t[10]-t[9]}  leaf+pow+store {t[11]-t[10]}  total {t[11]-t[8]} cycles")
    be.L.rlb_debug_set_tick_buffer(None)
