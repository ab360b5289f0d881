"""MPC Q-learning on the linear system -- the reference's usage example
(rlmpc/examples/linear_system_mpc_qlearning.py) on the B200 engine.

Same experiment: roll out the MPC policy for EPISODE_LENGTH steps in the noisy linear environment
(rlmpc/gym/linear_system/environment.py:6-66), then for every transition evaluate Q(s,a), dQ/dp and
V(s'), form the TD error ``cost + GAMMA * V(s') - Q(s,a)`` and move the parameters by
``mean_i(LR * td_i * dQ_dp_i)`` (example lines 171-205).  The reference does the learning part with a
Python loop of ``q_update`` / ``update`` calls, one acados solve each; here it is two batched engine
calls (``learn_batched``).  ``learn_loop`` keeps the reference's per-sample loop through the mirrored
``MPC`` API for comparison.  No gymnasium / stable-baselines3 dependency: the environment and the
episode buffer are the few numpy lines below.
"""
from __future__ import annotations

import numpy as np

N_EPISODES = 100
EPISODE_LENGTH = 100
GAMMA = 0.90
LR = 1e-4


class LinearSystemEnv:
    """rlmpc/gym/linear_system/environment.py:6-66 (true system differs from the MPC model)."""

    def __init__(self, min_observation, max_observation, lb_noise=-0.1, ub_noise=0.1, seed=0):
        self.A = np.array([[0.9, 0.35], [0.0, 1.1]])
        self.B = np.array([[0.0813], [0.2]])
        self.lo, self.hi = np.asarray(min_observation, float), np.asarray(max_observation, float)
        self.lb_noise, self.ub_noise = lb_noise, ub_noise
        self.rng = np.random.default_rng(seed)
        self.state = None

    def reset(self):
        self.state = np.array([0.5, 0.5])
        return self.state.copy()

    def step(self, action):
        action = np.asarray(action, float).reshape(-1)
        self.state = self.A @ self.state + self.B @ action + np.array([self.rng.uniform(self.lb_noise, self.ub_noise), 0.0])
        cost = 0.5 * self.state @ self.state + 0.5 * action @ action
        cost += 1e2 * float(np.any(self.lo - self.state > 0)) + 1e2 * float(np.any(self.state - self.hi > 0))
        return self.state.copy(), float(cost)


def rollout(mpc, env, episode_length=EPISODE_LENGTH):
    """One closed-loop episode with the MPC policy (example lines 156-167)."""
    obs = env.reset()
    mpc.reset(obs)
    S, A, C = [], [], []
    for _ in range(episode_length):
        action = mpc.get_action(obs)
        nxt, cost = env.step(action)
        S.append(obs); A.append(np.array(action, float).reshape(-1)); C.append(cost)
        obs = nxt
    return np.array(S), np.array(A), np.array(C)


def learn_loop(mpc, S, A, C, gamma=GAMMA, lr=LR):
    """The reference's learning step, sample by sample through the MPC API (example lines 171-205)."""
    n = S.shape[0] - 1
    dQ_dp = np.zeros((n, mpc.get_p().shape[0])); q = np.zeros(n); v = np.zeros(n)
    mpc.reset(S[0])
    for i in range(n):
        mpc.q_update(S[i], mpc.unscale_action(A[i]))
        dQ_dp[i, :] = mpc.get_dQ_dp()
        q[i] = mpc.get_Q()
        mpc.update(S[i])
        v[i] = mpc.get_V()
    cost = C[:n]
    td = cost[:-1] + gamma * v[1:] - q[:-1]
    dp = np.mean(np.vstack([lr * td[i] * dQ_dp[i, :] for i in range(n - 1)]), axis=0)
    return dp, td, q, v


def learn_batched(engine, mpc, S, A, C, gamma=GAMMA, lr=LR, max_sqp=100):
    """Same numbers from two batched calls: Q(s_i,a_i) with dQ/dp, and V(s_i)."""
    import torch

    n = S.shape[0] - 1
    dev = engine.device
    s = torch.tensor(S[:n], dtype=torch.float64, device=dev)
    a = torch.tensor(mpc.unscale_action(A[:n]), dtype=torch.float64, device=dev)
    engine.set_theta(mpc.get_p())
    engine.reset(s)
    oq = engine.solve_sens(s, a, max_sqp=max_sqp)
    engine.reset(s)
    _, v, stv = engine.solve(s, max_sqp=max_sqp)
    cost = torch.tensor(C[:n], dtype=torch.float64, device=dev)
    td = cost[:-1] + gamma * v[1:] - oq["cost"][:-1]
    ok = (oq["status"][:-1] == 0) & (stv[1:] == 0)
    # sum_i td_i dQ_i over valid samples on the device (what an all-reduce would combine across ranks)
    acc = engine.td_grad(td, oq["dL"][:-1].contiguous(), (~ok).to(torch.int32))
    ng = oq["dL"].shape[1]
    dp = (lr * acc[:ng] / (n - 1)).cpu().numpy()
    return dp, td.cpu().numpy(), oq["cost"].cpu().numpy(), v.cpu().numpy()


def main(n_episodes=N_EPISODES, episode_length=EPISODE_LENGTH, device=0, verbose=True):
    from mpc4rl_b200.mpc.linear_system.acados import AcadosMPC
    from mpc4rl_b200.problems import linear_system_param_nominal

    mpc = AcadosMPC(linear_system_param_nominal(), discount_factor=GAMMA, device=device)
    engine = mpc.batched(max_batch=episode_length, device=device)
    env = LinearSystemEnv(mpc.ocp_solver.acados_ocp.constraints.lbx, mpc.ocp_solver.acados_ocp.constraints.ubx)
    log = []
    for ep in range(n_episodes):
        S, A, C = rollout(mpc, env, episode_length)
        dp, td, _, _ = learn_batched(engine, mpc, S, A, C)
        log.append(dict(episode=ep, cost=float(C.sum()), td_error=float(td.mean()), p=mpc.get_parameter_values().copy()))
        if verbose:
            print(f"episode {ep}: total cost {C.sum():.3f}, mean TD error {td.mean():.4f}")
        mpc.set_parameter(mpc.get_parameter_values() + dp)
    return log


if __name__ == "__main__":
    main()
