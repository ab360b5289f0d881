"""MPC actor / critic modules for actor-critic training (TD3-style), SURVEY.md 8(f-1).

Batched, differentiable counterparts of the reference's ``Actor`` (rlmpc/td3/policies.py:125-222):
``forward`` evaluates the whole observation batch in one engine call instead of
``th.stack([self._predict(o) for o in obs])`` (:197), ``parameters()`` returns a real
``nn.Parameter`` (the reference builds detached tensors, :215-222), and gradients flow through the
NLP sensitivities.  No stable-baselines3 dependency: these are plain ``nn.Module``s with the same
method names (``forward``, ``_predict``), usable from any training loop.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from ..autograd import mpc_value_and_policy
from ..batched import BatchedMPC


class MPCActor(nn.Module):
    """pi_theta(s) = first input of the MPC, scaled to [-1, 1] like ``AcadosMPC.get_action`` of the
    cartpole (cartpole/acados.py:239-249, mpc.py:290-301)."""

    def __init__(self, engine: BatchedMPC, scale_action: bool = True, max_sqp: int = 50, warm_start: bool = False):
        super().__init__()
        self.engine = engine
        self.max_sqp = max_sqp
        self.warm_start = warm_start
        self.scale = scale_action
        self.theta = nn.Parameter(torch.tensor(engine.theta.reshape(-1)[: engine.ngrad], dtype=torch.float64))
        self.register_buffer("low", torch.tensor(engine.spec.lbu, dtype=torch.float64))
        self.register_buffer("high", torch.tensor(engine.spec.ubu, dtype=torch.float64))
        self.last_status = None

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        obs2 = obs.reshape(-1, self.engine.nx)
        u0, _, status = mpc_value_and_policy(self.theta, obs2, self.engine, max_sqp=self.max_sqp, reset=not self.warm_start)
        self.last_status = status
        if self.scale:
            low, high = self.low.to(u0.device), self.high.to(u0.device)
            u0 = 2.0 * ((u0 - low) / (high - low)) - 1.0
        return u0.to(obs.dtype) if obs.dtype.is_floating_point else u0

    def _predict(self, observation: torch.Tensor, deterministic: bool = True) -> torch.Tensor:
        with torch.no_grad():
            return self.forward(observation)


class MPCCritic(nn.Module):
    """Q_theta(s, a) = optimal cost with the first input clamped (mpc.py:52-96); action in [-1, 1] if
    ``scale_action``.  Shares ``theta`` with an actor when one is passed."""

    def __init__(self, engine: BatchedMPC, actor: MPCActor | None = None, scale_action: bool = True, max_sqp: int = 50):
        super().__init__()
        self.engine = engine
        self.max_sqp = max_sqp
        self.scale = scale_action
        self.theta = actor.theta if actor is not None else nn.Parameter(
            torch.tensor(engine.theta.reshape(-1)[: engine.ngrad], dtype=torch.float64))
        self.register_buffer("low", torch.tensor(engine.spec.lbu, dtype=torch.float64))
        self.register_buffer("high", torch.tensor(engine.spec.ubu, dtype=torch.float64))

    def forward(self, obs: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
        obs2 = obs.reshape(-1, self.engine.nx).to(torch.float64)
        a = action.reshape(-1, self.engine.nu).to(torch.float64)
        if self.scale:
            low, high = self.low.to(a.device), self.high.to(a.device)
            a = 0.5 * (high - low) * (a + 1.0) + low
        _, q, _ = mpc_value_and_policy(self.theta, obs2, self.engine, u0=a, max_sqp=self.max_sqp)
        return q

    def value(self, obs: torch.Tensor) -> torch.Tensor:
        _, v, _ = mpc_value_and_policy(self.theta, obs.reshape(-1, self.engine.nx).to(torch.float64), self.engine,
                                       max_sqp=self.max_sqp)
        return v
