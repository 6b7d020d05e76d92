"""rl_b200.data -- mirror of ``torchrl.data``'s replay-buffer surface for the B200 hot path."""
from .framestack import FrameStackStorage
from .checkpointers import StorageCheckpointerBase, TensorStorageCheckpointer
from .replay_buffers import (PrioritizedReplayBuffer, ReplayBuffer, TensorDictPrioritizedReplayBuffer,
                             TensorDictReplayBuffer)
from .samplers import (PrioritizedSampler, PrioritizedSliceSampler, RandomSampler, Sampler, SamplerWithoutReplacement,
                       SliceSampler, SliceSamplerWithoutReplacement)
from .segment_tree import (MinSegmentTreeFp32, MinSegmentTreeFp64, SumSegmentTreeFp32, SumSegmentTreeFp64)
from .storages import LazyTensorStorage, ListStorage, Storage, TensorStorage
from .tensordict_lite import TensorDict, is_tensor_collection
from .writers import RoundRobinWriter, TensorDictRoundRobinWriter, Writer

__all__ = [
    "ReplayBuffer", "PrioritizedReplayBuffer", "TensorDictReplayBuffer", "TensorDictPrioritizedReplayBuffer",
    "Sampler", "RandomSampler", "SamplerWithoutReplacement", "SliceSampler", "SliceSamplerWithoutReplacement", "PrioritizedSliceSampler", "PrioritizedSampler", "Storage", "ListStorage", "TensorStorage",
    "LazyTensorStorage", "FrameStackStorage", "Writer", "RoundRobinWriter", "TensorDictRoundRobinWriter", "TensorDict",
    "is_tensor_collection", "StorageCheckpointerBase", "TensorStorageCheckpointer", "SumSegmentTreeFp32", "SumSegmentTreeFp64", "MinSegmentTreeFp32", "MinSegmentTreeFp64",
]
