"""TEST INFRASTRUCTURE -- the reference's two CPU GAE code paths as torch-op sequences.

The reference's estimators are pure Python over torch ops; /root/reference cannot travel to the
GPU box, so this module restates the *same op sequences* for (a) timing the reference CPU path
there (bench.py cpu_baseline / --impl reference) and (b) cross-checking the C oracle.  It is
validated in the dev container against the imported reference itself
(tests/test_oracle.py::test_gae_torch_restatement_matches_reference, bit-for-bit on CPU).

  loop_gae  <- generalized_advantage_estimate        objectives/value/functional.py:119-180
  vec_gae   <- vec_generalized_advantage_estimate    functional.py:270-370 scalar-gamma branch
               -> _fast_vec_gae :211-267, _get_num_per_traj utils.py:195-210,
                  _split_and_pad_sequence utils.py:213-321, _geom_series_like functional.py:183-208,
                  _custom_conv1d utils.py:13-87 (2-D filter branch)

Inputs are [*B, T, F] with time at dim -2; done/terminated are bool.
"""
from __future__ import annotations

import math

import torch


def loop_gae(gamma, lmbda, state_value, next_state_value, reward, done, terminated=None):
    if terminated is None:
        terminated = done.clone()
    if not (next_state_value.shape == state_value.shape == reward.shape == done.shape == terminated.shape):
        raise RuntimeError("All input tensors (value, reward and done states) must share a unique shape.")
    not_done = (~done).int()
    not_term = (~terminated).int()
    T = not_done.shape[-2]
    adv = torch.empty_like(next_state_value)
    delta = reward + (gamma * not_term) * next_state_value - state_value
    disc = lmbda * gamma * not_done
    carry = 0
    for t in range(T - 1, -1, -1):
        carry = delta[..., t, :] + carry * disc[..., t, :]
        adv[..., t, :] = carry
    return adv, adv + state_value


def _traj_lengths(done_bt: torch.Tensor) -> torch.Tensor:
    # utils.py:195-210 -- the last step of every row closes a trajectory
    d = done_bt.clone()
    d[..., -1] = True
    ends = torch.where(d.reshape(-1))[0] + 1
    ends[1:] = ends[1:] - ends[:-1]
    return ends


def _pad_trajectories(x_bt: torch.Tensor, lengths: torch.Tensor):
    # utils.py:213-321 with return_mask=True on a plain tensor whose time dim is last
    tmax = torch.max(lengths)
    idt = torch.int16 if x_bt.size(-1) < torch.iinfo(torch.int16).max else torch.int32
    flat = x_bt.flatten(0, -1) if x_bt.ndim > 1 else x_bt
    steps = torch.arange(tmax, device=x_bt.device, dtype=idt).unsqueeze(0)
    mask = steps < lengths.to(x_bt.device).unsqueeze(1)
    padded = torch.zeros(len(lengths), tmax, dtype=x_bt.dtype, device=x_bt.device)
    padded = torch.masked_scatter(padded, mask, flat.reshape(-1))
    return padded, mask


def _geom_filter(like: torch.Tensor, ratio, thr: float) -> torch.Tensor:
    # functional.py:183-208 (non-compiled branch)
    if isinstance(ratio, torch.Tensor):
        ratio = ratio.item()
    if ratio == 0.0:
        return torch.zeros_like(like)
    lim = like.numel() if ratio >= 1.0 else int(math.log(thr) / math.log(ratio))
    rs = torch.full_like(like[:lim], ratio)
    rs[0] = 1.0
    return rs.cumprod(0).unsqueeze(-1)


def _one_sided_conv(x_b1t: torch.Tensor, filt_l1: torch.Tensor) -> torch.Tensor:
    # utils.py:77-81 -- right zero-pad by L-1 then a valid conv1d with the [1,1,L] filter
    padded = torch.nn.functional.pad(x_b1t, [0, filt_l1.shape[-2] - 1])
    return torch.conv1d(padded, filt_l1.squeeze(-1).unsqueeze(0).unsqueeze(0))


def vec_gae(gamma, lmbda, state_value, next_state_value, reward, done, terminated=None, thr: float = 1e-7):
    if terminated is None:
        terminated = done.clone()
    if not (next_state_value.shape == state_value.shape == reward.shape == done.shape == terminated.shape):
        raise RuntimeError("All input tensors (value, reward and done states) must share a unique shape.")
    # functional.py:243-247 -- time to the last dim
    d, tm, r, v, nv = (x.transpose(-2, -1) for x in (done, terminated, reward, state_value, next_state_value))
    gl = gamma * lmbda
    td0 = r + (~tm).int() * gamma * nv - v
    lengths = _traj_lengths(d)
    flat, mask = _pad_trajectories(td0, lengths)
    filt = _geom_filter(flat[0], gl, thr)
    adv = _one_sided_conv(flat.unsqueeze(1), filt).squeeze(1)
    adv = adv[mask].view_as(r)
    tgt = adv + v
    return adv.transpose(-1, -2), tgt.transpose(-1, -2)
