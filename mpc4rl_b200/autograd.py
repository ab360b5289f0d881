"""torch.autograd bridge: the batched MPC as a differentiable function of its parameters.

SURVEY.md 8(f-1).  The reference's TD3 actor calls ``mpc.get_action`` per observation in a Python
loop and returns detached tensors (rlmpc/td3/policies.py:186-213: "actor optimizer is not being
used", :332); its Q-learning drivers multiply TD errors with ``dQ_dp`` by hand
(examples/linear_system_mpc_qlearning.py:192-205).  Here one call evaluates the whole minibatch and
the NLP sensitivities the engine computes are the backward pass:

    d loss / d theta = sum_b ( dloss/du0_b . dpi_b/dtheta + dloss/dV_b . dL_b/dtheta )

``theta`` is the [ngrad] prefix of the reference's parameter vector p that has a gradient (the model
parameters; all of p for problems whose parameters are all "model", e.g. the linear system).
dV/dx0 is available too: the multiplier of the eliminated x_0 = s constraint (envelope theorem).
"""
from __future__ import annotations

from typing import Optional

import torch

from .batched import BatchedMPC


class _MPCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta: torch.Tensor, x0: torch.Tensor, u0: Optional[torch.Tensor], engine: BatchedMPC,
                max_sqp: int, reset: bool):
        # theta stays on the device: the full parameter vector p is kept as a CUDA tensor next to the engine and handed
        # over stream-ordered (rlmpc_set_theta_dev), no host round trip per call
        full = getattr(engine, "_theta_full_dev", None)
        if full is None:
            cur = engine.theta
            if cur.ndim != 1:
                raise ValueError("the autograd bridge needs one theta shared by the batch (engine.theta is per-sample)")
            full = torch.as_tensor(cur, dtype=torch.float64).to(engine.device).clone()
            engine._theta_full_dev = full
        full[: theta.numel()] = theta.detach().to(engine.device, torch.float64)
        engine.set_theta(full)
        x0d = x0.detach().to(engine.device, torch.float64).contiguous()
        u0d = None if u0 is None else u0.detach().to(engine.device, torch.float64).contiguous()
        if reset:
            engine.reset(x0d)
        out = engine.solve_sens(x0d, u0d, max_sqp=max_sqp)
        B = x0d.shape[0]
        ok = (out["status"] == 0)
        rho_x0 = engine.get("rho_x0", 0, B)
        # Q-mode: dQ/du0 is the multiplier of the clamped u_0 = a row (what makes critic(obs, actor(obs)) differentiable
        # through the action, the deterministic-policy-gradient use)
        rho_u0 = engine.get("rho_u0", 0, B) if u0 is not None else torch.zeros(B, engine.nu, dtype=torch.float64, device=engine.device)
        ctx.save_for_backward(out["dL"], out["dpi"], ok, rho_x0, rho_u0)
        ctx.theta_meta = (theta.device, theta.dtype, theta.numel(), x0.device, x0.dtype, x0.requires_grad,
                          None if u0 is None else (u0.device, u0.dtype, u0.requires_grad))
        ctx.mark_non_differentiable(out["status"])
        return out["u0"].to(x0.device), out["cost"].to(x0.device), out["status"].to(x0.device)

    @staticmethod
    def backward(ctx, g_u, g_v, _g_status):
        dL, dpi, ok, rho_x0, rho_u0 = ctx.saved_tensors
        tdev, tdt, nth, xdev, xdt, x_req, u_meta = ctx.theta_meta
        okf = ok.to(torch.float64)
        g = torch.zeros(dL.shape[1], dtype=torch.float64, device=dL.device)
        if g_v is not None:
            g = g + ((g_v.to(dL.device, torch.float64) * okf).unsqueeze(1) * dL).sum(0)
        if g_u is not None:
            gu = g_u.to(dL.device, torch.float64) * okf.unsqueeze(1)
            g = g + torch.einsum("bu,bup->p", gu, dpi)
        g_x0 = None
        if x_req and g_v is not None:
            g_x0 = ((g_v.to(dL.device, torch.float64) * okf).unsqueeze(1) * rho_x0).to(xdev, xdt)
        g_u0 = None
        if u_meta is not None and u_meta[2] and g_v is not None:
            g_u0 = ((g_v.to(dL.device, torch.float64) * okf).unsqueeze(1) * rho_u0).to(u_meta[0], u_meta[1])
        # (a loss on the returned u0 has no gradient w.r.t. x0 here: dpi/dx0 is not computed by the engine)
        return g[:nth].to(tdev, tdt), g_x0, g_u0, None, None, None


def mpc_value_and_policy(theta: torch.Tensor, x0: torch.Tensor, engine: BatchedMPC, u0: Optional[torch.Tensor] = None,
                         max_sqp: int = 50, reset: bool = True):
    """(u0 [B,nu], V or Q [B], status [B]) with gradients w.r.t. ``theta`` (and of V/Q w.r.t. ``x0``).
    Samples whose solve failed (status != 0) contribute zero gradient, like the reference's drivers that
    skip them (scripts/cartpole_mpc_qlearning_agent.py:167-180)."""
    return _MPCFunction.apply(theta, x0, u0, engine, int(max_sqp), bool(reset))
