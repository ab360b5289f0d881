"""Closed-loop MPC actor + MPC critic on vectorised cart-pole environments, all on the GPU
(BASELINE.json configs[4]).

Every environment step, for all environments at once:
  actor   a_t = pi_theta(s_t): one SQP-RTI step of the MPC from the environment's own warm start
          (the reference evaluates ``mpc.get_action`` per observation, rlmpc/td3/policies.py:197);
          the same call returns V_theta(s_t)
  env     (s_{t+1}, cost_t) = step(s_t, a_t)                      (rlmpc_cartpole_env_step)
  critic  Q_theta(s_{t-1}, a_{t-1}) and dQ/dtheta: MPC with the first input clamped (mpc.py:52-96)
  TD      td = cost_{t-1} + gamma V(s_t) - Q(s_{t-1}, a_{t-1});   theta += lr * mean(td * dQ/dtheta)
          (examples/linear_system_mpc_qlearning.py:192-205); the sum over environments is reduced on the
          device and, with several ranks, all-reduced (one small vector per step).
Run on N GPUs with ``python -m torch.distributed.run --nproc-per-node N -m mpc4rl_b200.examples.cartpole_mpc_actor_critic``.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ..batched import BatchedMPC
from ..gym.continuous_cartpole import ContinuousCartPoleSwingUpVectorEnv
from ..parallel import allreduce_accumulator
from ..problems import cartpole_original_config, cartpole_spec


def run(num_envs: int = 4096, n_steps: int = 50, gamma: float = 0.99, lr: float = 2e-3, device: int = 0,
        explore: float = 0.05, seed: int = 0, critic_sqp: int = 80, critic_step_length: float = 0.5, verbose: bool = True):
    """``lr`` is a RELATIVE step size: theta moves along mean_i(td_i dQ_i/dtheta) (the reference's semi-gradient
    Q-learning direction, examples/linear_system_mpc_qlearning.py:203) by at most ``lr`` of its own magnitude per
    environment step, and is projected onto physical values (masses and length within [0.2, 5] x nominal).  The
    reference uses a fixed absolute rate tuned by hand per problem (LR = 1e-4 for the linear system); with TD errors of
    the size the cart-pole produces that drives M, m, l negative within a few dozen steps."""
    spec = cartpole_spec(cartpole_original_config())
    dev = torch.device("cuda", device)
    actor = BatchedMPC(spec, max_batch=num_envs, device=device)    # owns one warm start per environment
    critic = BatchedMPC(spec, max_batch=num_envs, device=device)   # warm-started from the actor's iterate of the same state
    critic.set_option("tol", 1e-4)  # (update_nlp's own acceptance thresholds are 1e-3 / 1e-4, nlp.py:1513-1537)
    # full-step Gauss-Newton SQP 2-cycles on part of the swing-up (all environments pass through it together); a fixed
    # step length < 1 (acados: nlp_solver_step_length) converges there
    critic.set_option("step_length", critic_step_length)
    store = actor.iterate_store(num_envs)                          # the actor's iterate at s_{t-1}, one slot per environment
    slots = torch.arange(num_envs, dtype=torch.int32, device=dev)
    env = ContinuousCartPoleSwingUpVectorEnv(num_envs=num_envs, force_mag=float(spec.ubu[0]), device=device)
    rank = int(os.environ.get("RANK", "0"))
    g = torch.Generator(device="cpu").manual_seed(seed + 1000 * rank)
    ng = critic.ngrad
    theta = torch.tensor(spec.p_nominal, dtype=torch.float64, device=dev)  # stays on the device (rlmpc_set_theta_dev)
    th_nom = theta[:ng].clone()
    lo, hi = float(spec.lbu[0]), float(spec.ubu[0])

    s, _ = env.reset()
    actor.reset(s)
    actor.solve(s, max_sqp=30)  # converge the warm starts once
    prev = None
    log = []
    for t in range(n_steps):
        out = actor.solve_sens(s, max_sqp=1)  # RTI: pi(s_t), V(s_t)
        v_t, u_t = out["cost"], out["u0"]
        if prev is not None:
            # Q(s_{t-1}, a_{t-1}), warm-started from the actor's iterate at s_{t-1} (the V-problem there: the Q-problem
            # differs by the clamped first input only, so a few SQP iterations converge it)
            store.load_into(critic, slots)
            q = critic.solve_sens(prev["s"], prev["u"], max_sqp=critic_sqp)
            td = prev["cost"] + gamma * v_t * (~prev["term"]).double() - q["cost"]
            invalid = ((q["status"] == 0) & (out["status"] == 0) & ~prev["trunc"]).logical_not().to(torch.int32)
            acc = critic.td_grad(td, q["dL"], invalid)
            allreduce_accumulator(acc)  # the one collective of the step: [sum td dQ/dtheta (ng), sum td, n_valid]
            n_valid = acc[-1].clamp(min=1.0)
            gdir = acc[:ng] / n_valid
            rel = (gdir / th_nom).abs().max().clamp(min=1.0)  # at most `lr` relative change per step
            theta[:ng] = torch.minimum(torch.maximum(theta[:ng] + lr * th_nom * (gdir / th_nom) / rel, 0.2 * th_nom), 5.0 * th_nom)
            actor.set_theta(theta)
            critic.set_theta(theta)
            log.append(dict(step=t, mean_td=float((acc[-2] / n_valid).item()), n_valid=int(acc[-1].item()),
                            mean_cost=float(prev["cost"].mean().item()), theta=theta[:ng].cpu().numpy().copy(),
                            actor_bad=int((out["status"] != 0).sum().item()), critic_bad=int((q["status"] != 0).sum().item()),
                            critic_res=float(q["res"].max(dim=1).values.median().item()),
                            critic_codes=torch.bincount(q["status"].long(), minlength=5).tolist(),
                            actor_codes=torch.bincount(out["status"].long(), minlength=5).tolist()))
            if verbose and rank == 0:
                print(f"step {t}: mean cost {log[-1]['mean_cost']:.3f} mean TD {log[-1]['mean_td']:.3f} "
                      f"valid {log[-1]['n_valid']} (actor bad {log[-1]['actor_bad']}, critic bad {log[-1]['critic_bad']}, "
                      f"critic KKT median {log[-1]['critic_res']:.1e}; status codes actor {log[-1]['actor_codes']} critic "
                      f"{log[-1]['critic_codes']}) theta {log[-1]['theta']}", flush=True)
        store.save(slots)
        a = 2.0 * (u_t - lo) / (hi - lo) - 1.0
        a = (a + explore * torch.randn(a.shape, generator=g, dtype=torch.float64).to(dev)).clamp(-1.0, 1.0)
        u_applied = 0.5 * (hi - lo) * (a + 1.0) + lo
        s_next, cost, term, trunc, _ = env.step(a)
        prev = dict(s=s, u=u_applied, cost=cost, term=term, trunc=trunc)
        done = term | trunc
        if bool(done.any()):
            actor.reset(s_next, mask=done)  # like MPC.reset(obs) at the start of an episode
        s = s_next
    return log


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    run(num_envs=4096 // world, device=local, n_steps=int(os.environ.get("RLMPC_STEPS", "50")))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
