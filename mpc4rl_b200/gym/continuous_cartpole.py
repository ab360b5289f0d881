"""Vectorised continuous cart-pole swing-up environment on the GPU.

Same constructor arguments, dynamics, termination, auto-reset and reward as the reference's
``ContinuousCartPoleSwingUpVectorEnv`` (rlmpc/gym/continuous_cartpole/environment.py:302-459), but the
state lives in a CUDA tensor and ``step`` is one kernel launch (``rlmpc_cartpole_env_step``), so a
closed loop with the batched MPC never leaves the device.  No gymnasium dependency.

Two deliberate differences from the reference (both marked as unverified there, :416-419, :448-456):
the reward uses every environment's own action (the reference indexes ``action[0]``, i.e. environment 0's
action for all of them) and is evaluated on the state reached, before the auto-reset.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from .. import _cabi


class ContinuousCartPoleSwingUpVectorEnv:
    def __init__(self, num_envs: int = 2, max_episode_steps: int = 500, render_mode: Optional[str] = None,
                 min_action: float = -1.0, max_action: float = 1.0, force_mag: float = 30.0, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("the vectorised environment runs on CUDA devices only")
        self.lib = _cabi.load()
        self.device = torch.device("cuda", device)
        self.num_envs = int(num_envs)
        self.gravity, self.masscart, self.masspole, self.length = 9.8, 1.0, 0.1, 0.5
        self.force_mag, self.tau = float(force_mag), 0.02
        self.max_episode_steps = int(max_episode_steps)
        self.min_action, self.max_action = float(min_action), float(max_action)
        self.theta_threshold_radians = 360 * 2 * math.pi / 360
        self.x_threshold = 2.4
        self.reset_state = [0.0, 0.0, math.pi, 0.0]  # environment.py:441-442
        self._par = torch.tensor([self.gravity, self.masscart, self.masspole, self.length, self.force_mag, self.tau,
                                  self.x_threshold, self.theta_threshold_radians, float(self.max_episode_steps)]
                                 + self.reset_state, dtype=torch.float64, device=self.device)
        f64 = dict(dtype=torch.float64, device=self.device)
        i32 = dict(dtype=torch.int32, device=self.device)
        self.state = None
        self.steps = torch.zeros(self.num_envs, **i32)
        self._reward = torch.empty(self.num_envs, **f64)
        self._terminated = torch.empty(self.num_envs, **i32)
        self._truncated = torch.empty(self.num_envs, **i32)

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        self.state = torch.tensor(self.reset_state, dtype=torch.float64, device=self.device).repeat(self.num_envs, 1).contiguous()
        self.steps.zero_()
        return self.state.clone(), {}

    def step(self, action: torch.Tensor):
        assert self.state is not None, "Call reset before using step method."
        a = action.to(self.device, torch.float64).reshape(self.num_envs).contiguous()
        if not bool(((a >= self.min_action) & (a <= self.max_action)).all()):
            raise AssertionError("action outside the action space")
        p = lambda t: C.c_void_p(t.data_ptr())
        _cabi.check(self.lib.rlmpc_cartpole_env_step(p(self._par), self.num_envs, p(self.state), p(a), p(self._reward),
                                                     p(self._terminated), p(self._truncated), p(self.steps),
                                                     C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return self.state.clone(), self._reward.clone(), self._terminated.bool(), self._truncated.bool(), {}
