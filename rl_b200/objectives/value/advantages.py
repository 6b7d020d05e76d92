"""``GAE`` value-estimator module: mirror of ``torchrl.objectives.value.advantages.GAE`` (advantages.py:1338-1721).

Keeps the constructor keywords, ``set_keys``, ``forward(tensordict, *, params, target_params, time_dim)``,
``value_estimate`` and the key plumbing (reads ``("next", reward|done|terminated)`` and ``state_value``,
writes ``advantage`` / ``value_target``).  The estimator itself is one ``rlb_gae`` launch; the scalars
gamma and gamma*lmbda are rounded once at construction exactly as the reference's 0-d buffers would be, so
``forward`` never synchronises.  Calling the critic (``_call_value_nets``, advantages.py:476-619) is the
step before the hot path: a ``value_network`` callable is simply invoked on the tensordict and on its
``"next"`` sub-tensordict.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from ...data.tensordict_lite import is_tensor_collection
from .functional import (_gae_impl, gae_scalars, td0_return_estimate, vec_td1_return_estimate,
                         vec_td_lambda_return_estimate)


@dataclass
class _AcceptedKeys:
    """Default tensordict keys (advantages.py:113-144)."""

    advantage: object = "advantage"
    value_target: object = "value_target"
    value: object = "state_value"
    reward: object = "reward"
    done: object = "done"
    terminated: object = "terminated"
    steps_to_next_obs: object = "steps_to_next_obs"


def _nk(prefix, key) -> tuple:
    return (prefix, *key) if isinstance(key, tuple) else (prefix, key)


class GAE(nn.Module):
    """A class wrapper around the generalized advantage estimate functional.

    Refer to "HIGH-DIMENSIONAL CONTINUOUS CONTROL USING GENERALIZED ADVANTAGE ESTIMATION"
    https://arxiv.org/pdf/1506.02438.pdf for more context.

    Keyword Args:
        gamma (scalar): exponential mean discount.
        lmbda (scalar): trajectory discount.
        value_network (callable, optional): critic; called as ``value_network(td)`` and must write the value key.
            ``None``: values are read from the tensordict (``state_value`` and ``("next", "state_value")``).
        average_gae (bool): if ``True``, the resulting GAE values are standardized. Default ``False``.
        differentiable (bool): must stay ``False`` (the kernel is forward-only).
        vectorized (bool, optional): accepted; both settings run the same kernel.
        skip_existing (bool, optional): skip the computation when the output keys are already present.
        advantage_key / value_target_key / value_key: key overrides (deprecated in the reference; use set_keys).
        shifted (bool): accepted; only affects how a critic would be called.
        device: device of the gamma / lmbda buffers.
        time_dim (int, optional): time dimension of the input tensordict (default: last batch dim).
        auto_reset_env (bool): bootstrap truncated steps with ``gamma * value`` (advantages.py:1617-1620).
    """

    def __init__(self, *, gamma, lmbda, value_network=None, average_gae: bool = False, differentiable: bool = False,
                 vectorized: bool | None = None, skip_existing: bool | None = None, advantage_key=None,
                 value_target_key=None, value_key=None, shifted: bool = False, device=None, time_dim: int | None = None,
                 auto_reset_env: bool = False, deactivate_vmap: bool = False, value_chunk_size: int | None = None):
        super().__init__()
        if differentiable:
            raise NotImplementedError("GAE(differentiable=True) needs a backward pass; the B200 kernel is forward-only.")
        self.value_network = value_network
        self.differentiable = differentiable
        self.skip_existing = skip_existing
        self.shifted = shifted
        self.average_gae = average_gae
        self.vectorized = vectorized
        self.time_dim = time_dim
        self.auto_reset_env = auto_reset_env
        self.tensor_keys = _AcceptedKeys()
        self.set_keys(advantage=advantage_key, value_target=value_target_key, value=value_key)
        as_t = lambda x: x.detach().clone().to(device) if isinstance(x, torch.Tensor) else torch.tensor(x, device=device)
        self.register_buffer("gamma", as_t(gamma))
        self.register_buffer("lmbda", as_t(lmbda))
        self._scalar_cache: dict = {}

    def set_keys(self, **kwargs) -> None:
        for k, v in kwargs.items():
            if v is None:
                continue
            if not hasattr(self.tensor_keys, k):
                raise KeyError(f"{k} is not an accepted tensordict key for advantages")
            setattr(self.tensor_keys, k, v)

    @property
    def in_keys(self) -> list:
        tk = self.tensor_keys
        keys = [_nk("next", tk.reward), _nk("next", tk.done), _nk("next", tk.terminated)]
        if self.value_network is None:
            keys += [tk.value, _nk("next", tk.value)]
        return keys

    @property
    def out_keys(self) -> list:
        return [self.tensor_keys.advantage, self.tensor_keys.value_target]

    def _scalars(self, dtype: torch.dtype):
        # one host read of the 0-d buffers per dtype; forward() itself stays sync-free afterwards.  The cache is keyed on
        # the buffers' version counters and identities, so load_state_dict / .fill_() / annealing the discount (all
        # in-place or re-assignments) are picked up, as in the reference, which reads the buffers on every call
        key = (dtype, id(self.gamma), self.gamma._version, id(self.lmbda), self.lmbda._version)
        if self._scalar_cache.get("key") != key:
            self._scalar_cache = {"key": key, "value": gae_scalars(self.gamma.cpu(), self.lmbda.cpu(), dtype)}
        return self._scalar_cache["value"]

    def _get_time_dim(self, time_dim, data) -> int:
        # index of the time dimension among the tensordict's batch dims; the last one by default
        # (advantages.py:416-430)
        for cand in (time_dim, self.time_dim):
            if cand is not None:
                return data.ndim + cand if cand < 0 else cand
        return data.ndim - 1

    def _values(self, tensordict, params=None, target_params=None):
        tk = self.tensor_keys
        if self.value_network is None:
            value = tensordict.get(tk.value, None)
            next_value = tensordict.get(_nk("next", tk.value), None)
            if value is None:
                raise ValueError(f"The tensor with key {tk.value} is missing, and no value network was provided.")
            if next_value is None:
                raise ValueError(
                    f"The tensor with key {_nk('next', tk.value)} is missing, and no value network was provided.")
            return value, next_value
        with torch.no_grad():
            self.value_network(tensordict)
            value = tensordict.get(tk.value)
            nxt = tensordict.get("next")
            self.value_network(nxt)
            next_value = nxt.get(tk.value)
        return value, next_value

    @torch.no_grad()
    def forward(self, tensordict, *, params=None, target_params=None, time_dim: int | None = None):
        """Computes the GAE given the data in tensordict and writes ``advantage`` and ``value_target`` into it."""
        if not is_tensor_collection(tensordict):
            raise TypeError("GAE.forward expects a TensorDict-like input")
        if tensordict.batch_dims < 1:
            raise RuntimeError(
                "Expected input tensordict to have at least one dimension, got "
                f"tensordict.batch_size = {tensordict.batch_size}")
        tk = self.tensor_keys
        if self.skip_existing and tensordict.get(tk.advantage, None) is not None \
                and tensordict.get(tk.value_target, None) is not None:
            return tensordict
        reward = tensordict.get(_nk("next", tk.reward))
        steps = tensordict.get(tk.steps_to_next_obs, None)
        value, next_value = self._values(tensordict, params, target_params)
        done = tensordict.get(_nk("next", tk.done))
        terminated = tensordict.get(_nk("next", tk.terminated), None)
        if terminated is None:
            terminated = done
        gamma = None
        if steps is not None:
            # n-step transitions: gamma ** steps_to_next_obs, one discount per step (advantages.py:1576-1578) -- computed
            # BEFORE the auto-reset bootstrap below, which uses it too (:1615-1618)
            gamma = self.gamma.to(reward.device) ** steps.view_as(reward)
        if self.auto_reset_env:
            truncated = tensordict.get(("next", "truncated"))
            reward = reward + (self.gamma.to(reward.device) if gamma is None else gamma) * value * truncated
            terminated = done
        td = self._get_time_dim(time_dim, tensordict)
        if steps is not None:
            # the per-step form of the scan (rlb_affine_scan)
            adv, value_target = _gae_impl(gamma, self.lmbda.to(reward.device), value, next_value, reward, done,
                                          terminated, td)
        else:
            adv, value_target = _gae_impl(None, None, value, next_value, reward, done, terminated, td,
                                          scalars=self._scalars(value.dtype))
        if self.average_gae:
            loc = adv.mean()
            scale = adv.std().clamp_min(1e-4)
            adv = (adv - loc) / scale
        tensordict.set(tk.advantage, adv)
        tensordict.set(tk.value_target, value_target)
        return tensordict

    @torch.no_grad()
    def value_estimate(self, tensordict, params=None, target_params=None, time_dim: int | None = None, **kwargs):
        if tensordict.batch_dims < 1:
            raise RuntimeError(
                "Expected input tensordict to have at least one dimensions, got"
                f"tensordict.batch_size = {tensordict.batch_size}")
        tk = self.tensor_keys
        reward = tensordict.get(_nk("next", tk.reward))
        value, next_value = self._values(tensordict, params, target_params)
        done = tensordict.get(_nk("next", tk.done))
        terminated = tensordict.get(_nk("next", tk.terminated), None)
        td = self._get_time_dim(time_dim, tensordict)
        _, value_target = _gae_impl(None, None, value, next_value, reward, done,
                                    done if terminated is None else terminated, td,
                                    scalars=self._scalars(value.dtype))
        return value_target


class _ReturnEstimator(GAE):
    """Shared plumbing of the TD(0) / TD(1) / TD(lambda) estimator modules (advantages.py:622-1336): same keys, same
    ``forward`` (``value_target = value_estimate(...)``, ``advantage = value_target - value``), a different return."""

    def __init__(self, *, gamma, lmbda=1.0, value_network=None, average_rewards: bool = False,
                 differentiable: bool = False, vectorized: bool | None = None, skip_existing: bool | None = None,
                 advantage_key=None, value_target_key=None, value_key=None, shifted: bool = False, device=None,
                 time_dim: int | None = None, deactivate_vmap: bool = False, value_chunk_size: int | None = None):
        super().__init__(gamma=gamma, lmbda=lmbda, value_network=value_network, differentiable=differentiable,
                         vectorized=vectorized, skip_existing=skip_existing, advantage_key=advantage_key,
                         value_target_key=value_target_key, value_key=value_key, shifted=shifted, device=device,
                         time_dim=time_dim)
        self.average_rewards = average_rewards

    def _return(self, gamma, next_value, reward, done, terminated, time_dim):
        raise NotImplementedError

    @torch.no_grad()
    def value_estimate(self, tensordict, params=None, target_params=None, next_value=None, time_dim: int | None = None,
                       **kwargs):
        tk = self.tensor_keys
        reward = tensordict.get(_nk("next", tk.reward))
        gamma = self.gamma.to(reward.device)
        steps = tensordict.get(tk.steps_to_next_obs, None)
        if steps is not None:
            gamma = gamma ** steps.view_as(reward)
        else:
            gamma = float(self._scalars(reward.dtype)[0])
        if self.average_rewards:
            reward = reward - reward.mean()
            reward = reward / reward.std().clamp_min(self._reward_std_floor)
            tensordict.set(_nk("next", tk.reward), reward)    # the rewards are updated in place, like the reference
        if next_value is None:
            _, next_value = self._values(tensordict, params, target_params)
        done = tensordict.get(_nk("next", tk.done))
        terminated = tensordict.get(_nk("next", tk.terminated), None)
        if terminated is None:
            terminated = done
        return self._return(gamma, next_value, reward, done, terminated, self._get_time_dim(time_dim, tensordict))

    _reward_std_floor = 1e-4

    @torch.no_grad()
    def forward(self, tensordict, *, params=None, target_params=None, time_dim: int | None = None):
        if not is_tensor_collection(tensordict):
            raise TypeError(f"{type(self).__name__}.forward expects a TensorDict-like input")
        if tensordict.batch_dims < 1:
            raise RuntimeError("Expected input tensordict to have at least one dimensions, got"
                               f"tensordict.batch_size = {tensordict.batch_size}")
        tk = self.tensor_keys
        if self.skip_existing and tensordict.get(tk.advantage, None) is not None \
                and tensordict.get(tk.value_target, None) is not None:
            return tensordict
        value, next_value = self._values(tensordict, params, target_params)
        value_target = self.value_estimate(tensordict, next_value=next_value, time_dim=time_dim)
        tensordict.set(tk.advantage, value_target - value)
        tensordict.set(tk.value_target, value_target)
        return tensordict


class TD0Estimator(_ReturnEstimator):
    """Temporal Difference (TD(0)) estimate of advantage function, AKA bootstrapped temporal difference or 1-step return
    (advantages.py:622-841): ``value_target = reward + gamma * not_terminated * next_value`` (elementwise)."""

    _reward_std_floor = 1e-5

    def __init__(self, *, gamma, value_network=None, **kwargs):
        super().__init__(gamma=gamma, value_network=value_network, **kwargs)

    def _return(self, gamma, next_value, reward, done, terminated, time_dim):
        return td0_return_estimate(gamma=gamma, next_state_value=next_value, reward=reward, terminated=terminated,
                                   done=done)


class TD1Estimator(_ReturnEstimator):
    r""":math:`\infty`-Temporal Difference (TD(1)) estimate of advantage function (advantages.py:844-1071): the
    discounted return bootstrapped at trajectory ends -- the TD(lambda) scan kernel with lambda = 1."""

    def __init__(self, *, gamma, value_network=None, **kwargs):
        super().__init__(gamma=gamma, value_network=value_network, **kwargs)

    def _return(self, gamma, next_value, reward, done, terminated, time_dim):
        if isinstance(gamma, torch.Tensor):
            raise NotImplementedError("TD1Estimator with steps_to_next_obs (per-step gamma) is not supported")
        return vec_td1_return_estimate(gamma, next_value, reward, done=done, terminated=terminated, time_dim=time_dim)


class TDLambdaEstimator(_ReturnEstimator):
    r"""TD(:math:`\lambda`) estimate of advantage function (advantages.py:1074-1336): one ``rlb_td_lambda_return``
    launch whichever ``vectorized`` says."""

    def __init__(self, *, gamma, lmbda, value_network=None, **kwargs):
        super().__init__(gamma=gamma, lmbda=lmbda, value_network=value_network, **kwargs)

    def _return(self, gamma, next_value, reward, done, terminated, time_dim):
        if isinstance(gamma, torch.Tensor):
            raise NotImplementedError("TDLambdaEstimator with steps_to_next_obs (per-step gamma) is not supported")
        return vec_td_lambda_return_estimate(gamma, self._lmbda_host(), next_value, reward, done=done, terminated=terminated,
                                             time_dim=time_dim)

    def _lmbda_host(self) -> float:
        key = ("lmbda", id(self.lmbda), self.lmbda._version)   # one host read per value of the buffer
        if self._scalar_cache.get("lmbda_key") != key:
            self._scalar_cache["lmbda_key"], self._scalar_cache["lmbda"] = key, float(self.lmbda.cpu())
        return self._scalar_cache["lmbda"]
