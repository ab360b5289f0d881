"""GPU tests of the consumers of the batched gradients: the autograd bridge,
the batched MPC actor / critic modules, the vectorised environment and the closed-loop example."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "cartpole_original.npz"))


def _engine(B, tol=1e-10):
    from mpc4rl_b200 import BatchedMPC, cartpole_original_config, cartpole_spec

    m = BatchedMPC(cartpole_spec(cartpole_original_config()), max_batch=B, device=0)
    m.set_option("tol", tol)
    return m


def test_autograd_matches_engine_sensitivities_and_finite_differences(golden):
    from mpc4rl_b200.autograd import mpc_value_and_policy

    ok = golden["status"][:, 0] == 0
    x0 = torch.tensor(golden["x0"][ok][:8], dtype=torch.float64, device="cuda:0", requires_grad=True)
    eng = _engine(8)
    theta = torch.tensor(golden["theta"][:3], dtype=torch.float64, requires_grad=True)
    u0, V, st = mpc_value_and_policy(theta, x0, eng, max_sqp=200)
    assert (st == 0).all()
    w_u = torch.linspace(0.5, 1.5, 8, dtype=torch.float64, device="cuda:0").unsqueeze(1)
    loss = V.sum() + (w_u * u0).sum()
    loss.backward()
    g_ref = golden["dV"][ok][:8, :3].sum(0) + (w_u.cpu().numpy()[:, :, None] * golden["dpi"][ok][:8, :, :3]).sum((0, 1))
    assert np.abs(theta.grad.numpy() - g_ref).max() < 1e-6 * np.abs(g_ref).max()
    # dV/dx0 (multiplier of the eliminated x_0 = s constraint) against central differences of V
    gx = x0.grad.cpu().numpy()
    d = 1e-6
    fd = np.zeros_like(gx)
    for j in range(4):
        for sgn in (+1, -1):
            xp = x0.detach().clone()
            xp[:, j] += sgn * d
            _, Vp, stp = mpc_value_and_policy(theta.detach(), xp, eng, max_sqp=200)
            assert (stp == 0).all()
            fd[:, j] += sgn * Vp.cpu().numpy() / (2 * d)
    assert np.abs(fd - gx).max() < 1e-4 * max(1.0, np.abs(gx).max())


def test_actor_and_critic_modules(golden):
    from mpc4rl_b200.mpc.cartpole.acados import AcadosMPC
    from mpc4rl_b200 import cartpole_original_config
    from mpc4rl_b200.td3.policies import MPCActor, MPCCritic

    ok = golden["status"][:, 0] == 0
    obs = torch.tensor(golden["x0"][ok][:6], dtype=torch.float32, device="cuda:0")
    eng = _engine(6, tol=1e-8)
    actor = MPCActor(eng, max_sqp=200)
    assert [p.shape for p in actor.parameters()] == [torch.Size([3])]  # a real nn.Parameter
    a = actor(obs)
    assert a.shape == (6, 1) and a.dtype == torch.float32 and float(a.abs().max()) <= 1.0 + 1e-6
    # equals the reference-style per-observation loop through the mirrored AcadosMPC.get_action
    mpc = AcadosMPC(config=cartpole_original_config(), build=True)
    mpc.ocp_solver.engine.set_option("tol", 1e-8)
    for i in range(6):
        x = golden["x0"][ok][i].astype(np.float32).astype(np.float64)
        mpc.reset(x)
        ai = mpc.get_action(x)
        assert abs(float(a[i, 0]) - float(ai[0])) < 1e-5
    # gradients reach theta through the policy and through Q
    critic = MPCCritic(_engine(6, tol=1e-8), actor=actor, max_sqp=200)
    assert critic.theta is actor.theta
    q = critic(obs, a.detach())
    v = critic.value(obs)
    assert (q >= v - 1e-6 * v.abs()).all()  # Q(s,a) >= min_a Q(s,a) = V(s)
    (q.mean() + a.double().sum()).backward()
    assert actor.theta.grad is not None and torch.isfinite(actor.theta.grad).all() and actor.theta.grad.abs().max() > 0
    opt = torch.optim.SGD(actor.parameters(), lr=1e-9)
    before = actor.theta.detach().clone()
    opt.step()
    assert not torch.equal(before, actor.theta.detach())
    _ = actor(obs)
    assert np.allclose(eng.theta[:3], actor.theta.detach().numpy())  # the engine sees the updated parameters


def test_vector_env_matches_reference_dynamics():
    from mpc4rl_b200.gym.continuous_cartpole import ContinuousCartPoleSwingUpVectorEnv

    n = 257
    env = ContinuousCartPoleSwingUpVectorEnv(num_envs=n, force_mag=30.0, max_episode_steps=7)
    s, _ = env.reset()
    assert s.shape == (n, 4) and torch.allclose(s[:, 2], torch.full((n,), math.pi, dtype=torch.float64, device="cuda:0"))
    g = torch.Generator().manual_seed(0)
    state = s.cpu().numpy().T.copy()  # the reference keeps (4, n)
    for t in range(9):
        a = (2 * torch.rand(n, 1, generator=g, dtype=torch.float64) - 1)
        s, r, term, trunc, _ = env.step(a.cuda())
        # the reference's step (environment.py:372-426), numpy
        x, x_dot, theta, theta_dot = state
        force = a.numpy()[:, 0] * 30.0
        ct, st_ = np.cos(theta), np.sin(theta)
        temp = (force + 0.05 * theta_dot**2 * st_) / 1.1
        thetaacc = (9.8 * st_ - ct * temp) / (0.5 * (4.0 / 3.0 - 0.1 * ct**2 / 1.1))
        xacc = temp - 0.05 * thetaacc * ct / 1.1
        state = np.stack((x + 0.02 * x_dot, x_dot + 0.02 * xacc, theta + 0.02 * theta_dot, theta_dot + 0.02 * thetaacc))
        ang = ((state[2] + np.pi) % (2 * np.pi)) - np.pi
        rew = 2 * state[0]**2 + 0.01 * state[1]**2 + 2 * ang**2 + 0.01 * state[3]**2 + 0.001 * a.numpy()[:, 0]**2
        assert np.abs(r.cpu().numpy() - rew).max() < 1e-12
        done = (np.abs(state[0]) > 2.4) | (np.abs(state[2]) > 2 * np.pi) | ((t + 1) % 7 == 0)
        assert np.array_equal((term | trunc).cpu().numpy(), done)
        state[:, done] = np.array([0.0, 0.0, np.pi, 0.0])[:, None]
        assert np.abs(s.cpu().numpy().T - state).max() < 1e-12
    with pytest.raises(AssertionError):
        env.step(torch.full((n, 1), 1.5))


def test_closed_loop_actor_critic_example_runs():
    from mpc4rl_b200.examples.cartpole_mpc_actor_critic import run

    log = run(num_envs=96, n_steps=24, verbose=False)
    assert len(log) == 23
    # every environment contributes (the critic is warm-started from the actor's iterate and uses a damped SQP step)
    assert all(np.isfinite(l["mean_td"]) and l["n_valid"] >= 0.9 * 96 for l in log), [l["n_valid"] for l in log]
    th = np.array([l["theta"] for l in log])
    assert not np.array_equal(th[0], th[-1])
    assert (th > 0.2 * th[0] - 1e-12).all() and (th < 5.0 * th[0] + 1e-12).all()  # projected onto physical values
    assert np.abs(th[-1] / th[0] - 1.0).max() < 24 * 2e-3 + 1e-9                      # relative step size bound
    assert log[-1]["mean_cost"] < log[0]["mean_cost"]                                # the swing-up makes progress


def test_iterate_store_keeps_rti_accurate(golden):
    """warm starts stored per replay-buffer entry.  A minibatch sampled in a different
    order gets its own iterates back, so one RTI step reproduces the converged solution, while the same RTI
    step from the wrong (permuted) warm starts does not."""
    ok = np.where(golden["status"][:, 0] == 0)[0][:16]
    x = torch.tensor(golden["x0"][ok], dtype=torch.float64, device="cuda:0")
    eng = _engine(16)
    store = eng.iterate_store(capacity=100)
    slots = torch.arange(16, device="cuda:0", dtype=torch.int32) * 5 + 3
    eng.reset(x)
    u_conv = eng.solve(x, max_sqp=200)[0]
    store.save(slots)
    perm = torch.randperm(16, generator=torch.Generator().manual_seed(1)).cuda()
    # minibatch = the same entries in another order; engine still holds the iterates in the OLD order
    u_wrong = eng.solve(x[perm], max_sqp=1)[0]
    store.load(slots[perm])
    out = eng.solve_sens(x[perm], max_sqp=1)
    assert (out["status"] == 0).all()
    assert (out["u0"] - u_conv[perm]).abs().max().item() < 1e-8
    assert out["res"].max().item() < 1e-8
    assert (u_wrong - u_conv[perm]).abs().max().item() > 1e-3
    # out-of-range slots are skipped
    store.load(torch.full((16,), -1, dtype=torch.int32))
