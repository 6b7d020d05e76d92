"""rl_b200 -- a B200-native replay-and-advantage engine behind the TorchRL API.

    rl_b200.data                 ReplayBuffer / Sampler / Storage / Writer mirror of torchrl.data
    rl_b200.objectives.value     GAE + functional estimators mirror of torchrl.objectives.value
    rl_b200.ops                  torch-tensor level entry to the C ABI (include/rlb200.h, librlb200.so)

The compute is hand-written sm_100a CUDA (rl_b200/csrc) behind a plain C ABI; PyTorch provides device
memory, streams and torch.distributed only.  There is no CPU fallback.
"""
from . import ops  # noqa: F401

__version__ = "0.1.0"
