"""Small helpers shared by the replay-buffer mirror (reference: torchrl/data/replay_buffers/utils.py:1-120)."""
from __future__ import annotations

import numbers

import numpy as np
import torch

INT_CLASSES = (int, np.integer)


def _is_int(index) -> bool:
    """True for Python/numpy integers and 0-d integer arrays/tensors (utils.py `_is_int`)."""
    if isinstance(index, INT_CLASSES):
        return True
    if isinstance(index, (np.ndarray, torch.Tensor)):
        return index.ndim == 0
    return False


def _to_torch(data, device=None) -> torch.Tensor:
    if isinstance(data, np.generic):
        data = data.item()
    if not isinstance(data, torch.Tensor):
        data = torch.as_tensor(data)
    return data.to(device) if device is not None else data


def _reduce(tensor: torch.Tensor, reduction: str, dim=None):
    """Reduce a priority tensor the way the reference's `_reduce` does (utils.py:94-112)."""
    if reduction == "max":
        return tensor.max().item() if dim is None else tensor.max(dim=dim)[0]
    if reduction == "min":
        return tensor.min().item() if dim is None else tensor.min(dim=dim)[0]
    if reduction == "mean":
        return tensor.mean().item() if dim is None else tensor.mean(dim=dim)
    if reduction == "median":
        return tensor.median().item() if dim is None else tensor.median(dim=dim)[0]
    raise NotImplementedError(f"Unknown reduction method {reduction}")


def unravel_index(index: torch.Tensor, shape) -> tuple:
    """Flat index -> tuple of per-dimension indices (row-major), like torch.unravel_index."""
    out = []
    for dim in reversed(tuple(shape)):
        out.append(index % dim)
        index = torch.div(index, dim, rounding_mode="floor")
    return tuple(reversed(out))


def is_number(x) -> bool:
    return isinstance(x, numbers.Number)
