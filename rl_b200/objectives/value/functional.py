"""Functional advantage estimators: mirror of ``torchrl.objectives.value.functional`` for GAE.

    generalized_advantage_estimate       functional.py:119-180   (python time loop in the reference)
    vec_generalized_advantage_estimate   functional.py:270-370   (pad + conv1d in the reference)
    td_lambda_return_estimate / vec_...  functional.py:790-899, 1056-1210   (+ the *_advantage_estimate forms)
    td1_return_estimate / vec_...        functional.py:464-570, 648-707     (TD(lambda) with lmbda = 1)
    td0_return_estimate / advantage      functional.py:378-457              (elementwise, no scan)
    vtrace_advantage_estimate            functional.py:1297-1382            (``rlb_affine_scan`` after a torch prologue)
    GAE with per-step gamma / lmbda      functional.py:317-370              (same: rolling products, no [T, T] tensor)
    reward2go                            functional.py:1385-1460            (same scan, c = gamma * not_done)

Both names resolve to the same single kernel launch (``rlb_gae``, csrc/gae.cu): a warp-level discounted
reverse scan that reads every input element once and writes every output element once, with no host
synchronisation.  Signatures, the ``time_dim`` convention, the shape check and its error message are the
reference's.  Forward only (the reference runs GAE under ``no_grad`` unless ``differentiable=True``).
"""
from __future__ import annotations

import math

import torch

from ... import ops

SHAPE_ERR = "All input tensors (value, reward and done states) must share a unique shape."

__all__ = [
    "generalized_advantage_estimate", "vec_generalized_advantage_estimate", "gae_scalars",
    "td0_return_estimate", "td0_advantage_estimate",
    "td1_return_estimate", "vec_td1_return_estimate", "td1_advantage_estimate", "vec_td1_advantage_estimate",
    "td_lambda_return_estimate", "vec_td_lambda_return_estimate", "td_lambda_advantage_estimate",
    "vec_td_lambda_advantage_estimate", "vtrace_advantage_estimate", "reward2go",
]


def gae_scalars(gamma, lmbda, dtype: torch.dtype) -> tuple[float, float]:
    """(gamma, gamma*lmbda) rounded the way the reference rounds them.

    Tensor arguments (what the GAE module passes, advantages.py:1456-1467) are multiplied as 0-d tensors
    (functional.py:249), so the product is rounded once in the tensors' dtype; Python numbers are multiplied
    in double and only cast when they meet a tensor.  Returned as Python floats (exactly representable).
    Calling this with CUDA tensors synchronises -- the GAE module caches the result instead.
    """
    if isinstance(gamma, torch.Tensor) or isinstance(lmbda, torch.Tensor):
        g = gamma if isinstance(gamma, torch.Tensor) else torch.tensor(gamma)
        l = lmbda if isinstance(lmbda, torch.Tensor) else torch.tensor(lmbda, device=g.device)
        if g.numel() > 1 or l.numel() > 1:
            raise ValueError("gae_scalars: per-step gamma / lmbda tensors have no scalar form")
        gl = (g * l.to(g.device)).to(dtype)
        return float(g.to(dtype)), float(gl)
    cast = torch.tensor([float(gamma), float(gamma) * float(lmbda)], dtype=dtype)
    return float(cast[0]), float(cast[1])


def _time_to_minus2(t: torch.Tensor, time_dim: int):
    """Bring ``time_dim`` to position -2 (functional.py:47-111); returns (tensor, squeeze_back)."""
    td = time_dim - t.ndim if time_dim >= 0 else time_dim
    if td < -t.ndim or td >= 0:
        raise RuntimeError(
            f"The tensor shape and the time dimension are not compatible: got {t.shape} and time_dim={td}.")
    if t.ndim >= 2:
        return t.transpose(td, -2), False
    if t.ndim == 1 and td == -1:
        return t.unsqueeze(-1), True
    raise RuntimeError(f"The tensor shape and the time dimension are not compatible: got {t.shape} and time_dim={td}.")


def _is_per_step(x) -> bool:
    return isinstance(x, torch.Tensor) and x.numel() > 1


def _check_forward_only(*tensors):
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise NotImplementedError(
            "the B200 advantage kernels are forward-only: call them under torch.no_grad() or detach the inputs.")


def _affine_scan(d: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """out_t = d_t + c_t * out_{t+1} along dim -2 of [*B, T, F] tensors (``rlb_affine_scan``, csrc/gae.cu MODE 2)."""
    shape = d.shape
    d = d.contiguous()
    c = c.to(d.dtype).expand(shape).contiguous()
    out = ops.backend().affine_scan(d, c, math.prod(shape[:-2]), shape[-2], shape[-1])
    return out.view(shape)


def _gae_per_step(gamma, lmbda, state_value, next_state_value, reward, done, terminated, time_dim):
    """GAE with tensor-valued gamma / lmbda (functional.py:317-370).

    The reference rolls ``not_done * gamma * lmbda`` into a [B*F, T, T+1, 1] tensor (value/utils.py:130-181), takes its
    cumprod and convolves; row t of that tensor is the running product of the coefficients from step t on, i.e. the
    recurrence  A_t = td0_t + (not_done_t * gamma_t * lmbda_t) * A_{t+1}, which is what the kernel scans.
    """
    dtype = state_value.dtype
    if dtype not in (torch.float32, torch.float64):
        raise NotImplementedError(f"GAE kernel supports fp32 / fp64 values, got {dtype}")
    tensors = [state_value, next_state_value, reward, done, terminated]
    coefs = [gamma, lmbda]
    nd = state_value.ndim
    td = time_dim - nd if time_dim >= 0 else time_dim
    squeeze = False
    if td != -2 or nd < 2:
        moved = [_time_to_minus2(t, time_dim) for t in tensors]
        squeeze = any(s for _, s in moved)
        tensors = [t for t, _ in moved]
        coefs = [_time_to_minus2(c, time_dim)[0] if _is_per_step(c) else c for c in coefs]
    v, nv, r, d, tm = (t.detach() for t in tensors)
    gamma, lmbda = (c.detach().to(v.device) if isinstance(c, torch.Tensor) else c for c in coefs)
    value = gamma * lmbda                                              # functional.py:317
    gammalmbdas = (~d.to(torch.bool)).to(dtype) * value                # :321-322
    td0 = r + (~tm.to(torch.bool)).to(dtype) * gamma * nv - v          # :349-350
    adv = _affine_scan(td0.to(dtype), gammalmbdas)
    tgt = adv + v                                                      # :369
    if squeeze:
        return adv.squeeze(-1), tgt.squeeze(-1)
    if td != -2:
        adv, tgt = adv.transpose(td, -2), tgt.transpose(td, -2)
    return adv, tgt


def _gae_impl(gamma, lmbda, state_value, next_state_value, reward, done, terminated, time_dim, scalars=None):
    if terminated is None:
        terminated = done
    if not (next_state_value.shape == state_value.shape == reward.shape == done.shape == terminated.shape):
        raise RuntimeError(SHAPE_ERR)
    if scalars is None and (_is_per_step(gamma) or _is_per_step(lmbda)):
        _check_forward_only(state_value, next_state_value, reward, gamma, lmbda)
        return _gae_per_step(gamma, lmbda, state_value, next_state_value, reward, done, terminated, time_dim)
    if state_value.requires_grad or next_state_value.requires_grad or reward.requires_grad:
        if torch.is_grad_enabled():
            raise NotImplementedError(
                "the B200 GAE kernel is forward-only: call it under torch.no_grad() (GAE(differentiable=False), the "
                "default) or detach the inputs.")
    dtype = state_value.dtype
    if dtype not in (torch.float32, torch.float64):
        raise NotImplementedError(f"GAE kernel supports fp32 / fp64 values, got {dtype}")
    squeeze = False
    tensors = [state_value, next_state_value, reward, done, terminated]
    nd = state_value.ndim
    td = time_dim - nd if time_dim >= 0 else time_dim  # normalised to a negative index
    if td != -2 or nd < 2:
        moved = [_time_to_minus2(t, time_dim) for t in tensors]
        squeeze = any(s for _, s in moved)
        tensors = [t for t, _ in moved]
    v, nv, r, d, tm = tensors
    shape = v.shape
    T, F = shape[-2], shape[-1]
    rows = math.prod(shape[:-2])
    v = v.detach().contiguous()
    nv = nv.detach().to(dtype).contiguous()
    r = r.detach().to(dtype).contiguous()
    d = d.to(torch.bool).contiguous().view(torch.uint8)
    tm = tm.to(torch.bool).contiguous().view(torch.uint8)
    g, gl = scalars if scalars is not None else gae_scalars(gamma, lmbda, dtype)
    adv, tgt = ops.backend().gae(v, nv, r, d, tm, g, gl, rows, T, F)
    adv, tgt = adv.view(shape), tgt.view(shape)
    if squeeze:
        return adv.squeeze(-1), tgt.squeeze(-1)
    if td != -2:
        adv, tgt = adv.transpose(td, -2), tgt.transpose(td, -2)
    return adv, tgt


def vec_generalized_advantage_estimate(gamma, lmbda, state_value: torch.Tensor, next_state_value: torch.Tensor,
                                       reward: torch.Tensor, done: torch.Tensor,
                                       terminated: torch.Tensor | None = None, *, time_dim: int = -2):
    """Vectorized Generalized advantage estimate of a trajectory (https://arxiv.org/pdf/1506.02438.pdf).

    Args:
        gamma (scalar or Tensor): exponential mean discount; a tensor shaped like ``done`` gives one value per step.
        lmbda (scalar or Tensor): trajectory discount; likewise.
        state_value (Tensor): value function result with old_state input.
        next_state_value (Tensor): value function result with new_state input.
        reward (Tensor): reward of taking actions in the environment.
        done (Tensor): boolean flag for end of trajectory.
        terminated (Tensor): boolean flag for the end of episode. Defaults to ``done`` if not provided.
        time_dim (int): dimension where the time is unrolled. Defaults to -2.

    All tensors (values, reward and done) must have shape ``[*Batch x TimeSteps x *F]``.
    Returns ``(advantage, value_target)``.
    """
    return _gae_impl(gamma, lmbda, state_value, next_state_value, reward, done, terminated, time_dim)


def generalized_advantage_estimate(gamma, lmbda, state_value: torch.Tensor, next_state_value: torch.Tensor,
                                   reward: torch.Tensor, done: torch.Tensor, terminated: torch.Tensor | None = None,
                                   *, time_dim: int = -2):
    """Generalized advantage estimate of a trajectory -- same kernel as the vectorized entry point (the
    reference's python time loop, functional.py:164-178, is what the kernel's recurrence restates)."""
    return _gae_impl(gamma, lmbda, state_value, next_state_value, reward, done, terminated, time_dim)


########################################################################
# TD(0), TD(1), TD(lambda)  -- SURVEY.md section 8(f)-2: the same reverse scan with other coefficients
# --------------------------------------------------------------------------------------------------


def td0_return_estimate(gamma, next_state_value: torch.Tensor, reward: torch.Tensor,
                        terminated: torch.Tensor | None = None, *, done: torch.Tensor | None = None) -> torch.Tensor:
    """TD(0) discounted return estimate ``r + gamma * (1 - terminated) * v'`` (functional.py:418-457).
    Elementwise -- there is nothing to scan, so it stays one fused torch expression as in the reference."""
    if done is not None and terminated is None:
        terminated = done.clone()
    if not (next_state_value.shape == reward.shape == terminated.shape):
        raise RuntimeError(SHAPE_ERR)
    return reward + gamma * (~terminated).int() * next_state_value


def td0_advantage_estimate(gamma, state_value, next_state_value, reward, done, terminated=None) -> torch.Tensor:
    """TD(0) advantage estimate (functional.py:378-415)."""
    if terminated is None:
        terminated = done.clone()
    if not (next_state_value.shape == state_value.shape == reward.shape == done.shape == terminated.shape):
        raise RuntimeError(SHAPE_ERR)
    return td0_return_estimate(gamma, next_state_value, reward, terminated) - state_value


def _td_lambda_scalars(gamma, lmbda, dtype):
    """(gamma, gamma*lmbda, 1-lmbda) rounded like the tensor ops of _fast_td_lambda_return_estimate
    (functional.py:1031-1046): every factor lives in the value dtype."""
    for x in (gamma, lmbda):
        if isinstance(x, torch.Tensor) and x.numel() > 1:
            raise NotImplementedError(
                "tensor-valued gamma / lmbda (one value per step) are not supported by the B200 scan kernel yet; "
                "pass scalars.")
    g = torch.as_tensor(gamma).detach().to("cpu", dtype).reshape(())
    l = torch.as_tensor(lmbda).detach().to("cpu", dtype).reshape(())
    return float(g), float(g * l), float(1 - l)


def _td_lambda_impl(gamma, lmbda, next_state_value, reward, done, terminated, rolling_gamma, time_dim):
    if terminated is None:
        terminated = done
    if not (next_state_value.shape == reward.shape == done.shape == terminated.shape):
        raise RuntimeError(SHAPE_ERR)
    if rolling_gamma is not None and not rolling_gamma:
        raise RuntimeError("rolling_gamma=False is expected only with time-sensitive gamma or lambda values")
    if (next_state_value.requires_grad or reward.requires_grad) and torch.is_grad_enabled():
        raise NotImplementedError("the B200 scan kernel is forward-only: call it under torch.no_grad().")
    dtype = next_state_value.dtype
    if dtype not in (torch.float32, torch.float64):
        raise NotImplementedError(f"TD(lambda) kernel supports fp32 / fp64 values, got {dtype}")
    squeeze = False
    tensors = [next_state_value, reward, done, terminated]
    nd = next_state_value.ndim
    td = time_dim - nd if time_dim >= 0 else time_dim
    if td != -2 or nd < 2:
        moved = [_time_to_minus2(t, time_dim) for t in tensors]
        squeeze = any(sq for _, sq in moved)
        tensors = [t for t, _ in moved]
    nv, r, d, tm = tensors
    shape = nv.shape
    T, F = shape[-2], shape[-1]
    rows = math.prod(shape[:-2])
    nv = nv.detach().contiguous()
    r = r.detach().to(dtype).contiguous()
    d = d.to(torch.bool).contiguous().view(torch.uint8)
    tm = tm.to(torch.bool).contiguous().view(torch.uint8)
    g, gl, oml = _td_lambda_scalars(gamma, lmbda, dtype)
    ret = ops.backend().td_lambda_return(nv, r, d, tm, g, gl, oml, rows, T, F).view(shape)
    if squeeze:
        return ret.squeeze(-1)
    if td != -2:
        ret = ret.transpose(td, -2)
    return ret


def vec_td_lambda_return_estimate(gamma, lmbda, next_state_value: torch.Tensor, reward: torch.Tensor,
                                  done: torch.Tensor, terminated: torch.Tensor | None = None,
                                  rolling_gamma: bool | None = None, *, time_dim: int = -2) -> torch.Tensor:
    r"""Vectorized TD(:math:`\lambda`) return estimate (functional.py:1056-1210; scalar ``gamma`` / ``lmbda``).

    Args:
        gamma (scalar): exponential mean discount.
        lmbda (scalar): trajectory discount.
        next_state_value (Tensor): value function result with new_state input.
        reward (Tensor): reward of taking actions in the environment.
        done (Tensor): boolean flag for end of trajectory.
        terminated (Tensor): boolean flag for the end of episode. Defaults to ``done`` if not provided.
        rolling_gamma (bool, optional): only meaningful for tensor-valued gamma; must be ``None`` / ``True`` here.
        time_dim (int): dimension where the time is unrolled. Defaults to -2.

    All tensors (values, reward and done) must have shape ``[*Batch x TimeSteps x *F]``.
    """
    return _td_lambda_impl(gamma, lmbda, next_state_value, reward, done, terminated, rolling_gamma, time_dim)


def td_lambda_return_estimate(gamma, lmbda, next_state_value, reward, done, terminated=None, rolling_gamma=None, *,
                              time_dim: int = -2) -> torch.Tensor:
    r"""TD(:math:`\lambda`) return estimate (functional.py:790-899) -- same kernel as the vectorized entry point."""
    return _td_lambda_impl(gamma, lmbda, next_state_value, reward, done, terminated, rolling_gamma, time_dim)


def _td_lambda_adv(gamma, lmbda, state_value, next_state_value, reward, done, terminated, rolling_gamma, time_dim):
    if terminated is None:
        terminated = done
    if not (next_state_value.shape == state_value.shape == reward.shape == done.shape == terminated.shape):
        raise RuntimeError(SHAPE_ERR)
    return _td_lambda_impl(gamma, lmbda, next_state_value, reward, done, terminated, rolling_gamma,
                           time_dim) - state_value


def vec_td_lambda_advantage_estimate(gamma, lmbda, state_value, next_state_value, reward, done, terminated=None,
                                     rolling_gamma=None, time_dim: int = -2) -> torch.Tensor:
    r"""Vectorized TD(:math:`\lambda`) advantage estimate: returns - state_value (functional.py:1213-1294)."""
    return _td_lambda_adv(gamma, lmbda, state_value, next_state_value, reward, done, terminated, rolling_gamma, time_dim)


def td_lambda_advantage_estimate(gamma, lmbda, state_value, next_state_value, reward, done, terminated=None,
                                 rolling_gamma=None, time_dim: int = -2) -> torch.Tensor:
    r"""TD(:math:`\lambda`) advantage estimate (functional.py:913-990)."""
    return _td_lambda_adv(gamma, lmbda, state_value, next_state_value, reward, done, terminated, rolling_gamma, time_dim)


def vec_td1_return_estimate(gamma, next_state_value, reward, done, terminated=None, rolling_gamma=None,
                            time_dim: int = -2) -> torch.Tensor:
    """Vectorized TD(1) return estimate = TD(lambda) with lmbda = 1 (functional.py:648-707)."""
    return _td_lambda_impl(gamma, 1, next_state_value, reward, done, terminated, rolling_gamma, time_dim)


def td1_return_estimate(gamma, next_state_value, reward, done, terminated=None, rolling_gamma=None,
                        time_dim: int = -2) -> torch.Tensor:
    """TD(1) return estimate (functional.py:464-570) -- same kernel as the vectorized entry point."""
    return _td_lambda_impl(gamma, 1, next_state_value, reward, done, terminated, rolling_gamma, time_dim)


def vec_td1_advantage_estimate(gamma, state_value, next_state_value, reward, done, terminated=None, rolling_gamma=None,
                               time_dim: int = -2) -> torch.Tensor:
    """Vectorized TD(1) advantage estimate (functional.py:710-787)."""
    return _td_lambda_adv(gamma, 1, state_value, next_state_value, reward, done, terminated, rolling_gamma, time_dim)


def td1_advantage_estimate(gamma, state_value, next_state_value, reward, done, terminated=None, rolling_gamma=None,
                           time_dim: int = -2) -> torch.Tensor:
    """TD(1) advantage estimate (functional.py:572-645)."""
    return _td_lambda_adv(gamma, 1, state_value, next_state_value, reward, done, terminated, rolling_gamma, time_dim)


########################################################################
# V-trace  -- SURVEY.md section 8(f)-2
# ------------------------------------


def vtrace_advantage_estimate(gamma, log_pi: torch.Tensor, log_mu: torch.Tensor, state_value: torch.Tensor,
                              next_state_value: torch.Tensor, reward: torch.Tensor, done: torch.Tensor,
                              terminated: torch.Tensor | None = None, rho_thresh=1.0, c_thresh=1.0,
                              time_dim: int = -2):
    """V-Trace off-policy actor-critic targets (IMPALA, https://arxiv.org/abs/1802.01561); functional.py:1297-1382.

    Args and shapes are the reference's: ``log_pi`` / ``log_mu`` the collection and current log-probabilities,
    ``rho_thresh`` / ``c_thresh`` the importance-weight clips, every tensor ``[*Batch x TimeSteps x *F]``.
    Returns ``(advantages, vs)``.  The reference's python loop over time (:1362-1368) is one ``rlb_affine_scan``
    launch here; the clipping prologue and the shifted epilogue stay elementwise torch ops with the reference's
    operation order.
    """
    if not (next_state_value.shape == state_value.shape == reward.shape == done.shape):
        raise RuntimeError(SHAPE_ERR)
    _check_forward_only(log_pi, log_mu, state_value, next_state_value, reward)
    dtype = state_value.dtype
    if dtype not in (torch.float32, torch.float64):
        raise NotImplementedError(f"V-trace kernel supports fp32 / fp64 values, got {dtype}")
    tensors = [log_pi, log_mu, state_value, next_state_value, reward, done, terminated]
    nd = state_value.ndim
    td = time_dim - nd if time_dim >= 0 else time_dim
    squeeze = False
    if td != -2 or nd < 2:
        moved = [_time_to_minus2(t, time_dim) if _is_per_step(t) else (t, False) for t in tensors]
        squeeze = any(s for _, s in moved)
        tensors = [t for t, _ in moved]
    log_pi, log_mu, v, nv, r, d, tm = (t.detach() if isinstance(t, torch.Tensor) else t for t in tensors)
    device = v.device
    rho_thresh = torch.as_tensor(rho_thresh, device=device)
    c_thresh = torch.as_tensor(c_thresh, device=device)
    not_done = (~d).int()                                              # :1347
    not_terminated = not_done if tm is None else (~tm).int()
    done_discounts = gamma * not_done
    terminated_discounts = gamma * not_terminated
    rho = (log_pi - log_mu).exp()
    clipped_rho = rho.clamp_max(rho_thresh)
    deltas = clipped_rho * (r + terminated_discounts * nv - v)         # :1355-1357
    clipped_c = rho.clamp_max(c_thresh)
    vs_minus_v = _affine_scan(deltas.to(dtype), done_discounts * clipped_c)   # :1360-1370
    vs = vs_minus_v + v
    vs_t_plus_1 = torch.cat([vs[..., 1:, :], nv[..., -1:, :]], dim=-2)
    advantages = clipped_rho * (r + terminated_discounts * vs_t_plus_1 - v)
    if squeeze:
        return advantages.squeeze(-1), vs.squeeze(-1)
    if td != -2:
        advantages, vs = advantages.transpose(td, -2), vs.transpose(td, -2)
    return advantages, vs


########################################################################
# Reward to go
# ------------


def reward2go(reward: torch.Tensor, done: torch.Tensor, gamma, *, time_dim: int = -2) -> torch.Tensor:
    """Discounted cumulative sum of rewards over trajectories delimited by ``done`` (functional.py:1385-1460).

    The reference splits the batch at ``done``, pads the pieces and convolves them with a geometric series
    (truncated below 1e-7); per element that is  G_t = r_t + gamma * (1 - done_t) * G_{t+1}  with the last step of the
    time axis closing its trajectory -- one ``rlb_affine_scan`` launch.  Same shape in, same shape out.
    """
    if reward.shape != done.shape:
        raise ValueError(f"reward and done must share the same shape, got {reward.shape} and {done.shape}")
    _check_forward_only(reward)
    dtype = reward.dtype
    if dtype not in (torch.float32, torch.float64):
        raise NotImplementedError(f"reward2go kernel supports fp32 / fp64 rewards, got {dtype}")
    nd = reward.ndim
    td = time_dim - nd if time_dim >= 0 else time_dim
    squeeze = False
    r, d = reward.detach(), done
    if td != -2 or nd < 2:
        (r, squeeze), (d, _) = _time_to_minus2(r, time_dim), _time_to_minus2(d, time_dim)
    out = _affine_scan(r, (~d.to(torch.bool)).to(dtype) * gamma)
    if squeeze:
        return out.squeeze(-1)
    return out.transpose(td, -2) if td != -2 else out
