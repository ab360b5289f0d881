"""GPU parity tests for the cartpole hot path (BASELINE.json configs[1]) -- all calls go through
the C ABI (librlmpc_b200.so).  Oracle = oracle/ (dense restatement); golden fixtures were produced
by oracle/make_golden.py.  Tolerances: |u0| 1e-6 abs (north-star: 1e-5),
V/Q 1e-9 rel, dQ/dtheta 1e-6 rel, dpi/dtheta 1e-5 rel."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "cartpole_original.npz"))


@pytest.fixture(scope="module")
def spec():
    from mpc4rl_b200 import cartpole_original_config, cartpole_spec

    return cartpole_spec(cartpole_original_config())


def _mpc(spec, B):
    from mpc4rl_b200 import BatchedMPC

    m = BatchedMPC(spec, max_batch=B, device=0)
    m.set_option("tol", 1e-10)  # the golden fixtures were converged to 1e-10 as well
    return m


def _dev(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64, device="cuda:0")


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def test_v_mode_matches_golden(spec, golden):
    x0 = _dev(golden["x0"])
    m = _mpc(spec, x0.shape[0])
    m.reset(x0)
    out = m.solve_sens(x0, max_sqp=200)
    ok = golden["status"][:, 0] == 0
    # full-step SQP (acados default) 2-cycles on two samples of this seed -- in the oracle too
    assert np.array_equal(out["status"].cpu().numpy() == 0, ok)
    golden = {k: golden[k][ok] for k in ("u0", "V", "dV", "dpi", "X", "U", "pi")}
    out = {k: v[torch.tensor(ok, device=v.device)] for k, v in out.items()}
    assert np.abs(out["u0"].cpu().numpy() - golden["u0"]).max() < 1e-6
    assert _rel(out["cost"].cpu().numpy(), golden["V"]) < 1e-9
    # the reference's p has 83 entries of which only (M, m, l) carry gradient (quirk Q8): the
    # engine returns that non-zero prefix, full_grad() pads to the reference layout
    assert out["dL"].shape[1] == 3 and np.count_nonzero(golden["dV"][:, 3:]) == 0
    assert np.count_nonzero(golden["dpi"][:, :, 3:]) == 0
    assert _rel(m.full_grad(out["dL"]).cpu().numpy(), golden["dV"]) < 1e-6
    assert _rel(m.full_grad(out["dpi"]).cpu().numpy(), golden["dpi"]) < 1e-5
    # full primal-dual solution
    N = spec.N
    X = np.stack([m.get("x", k, x0.shape[0]).cpu().numpy() for k in range(N + 1)], axis=1)[ok]
    U = np.stack([m.get("u", k, x0.shape[0]).cpu().numpy() for k in range(N)], axis=1)[ok]
    PI = np.stack([m.get("pi", k, x0.shape[0]).cpu().numpy() for k in range(N)], axis=1)[ok]
    assert np.abs(X - golden["X"]).max() < 2e-6
    assert np.abs(U - golden["U"]).max() < 2e-6
    assert np.abs(PI.reshape(len(PI), -1) - golden["pi"]).max() < 1e-5 * max(1.0, np.abs(golden["pi"]).max())


def test_q_mode_matches_golden(spec, golden):
    x0, a = _dev(golden["x0"]), _dev(golden["a"])
    m = _mpc(spec, x0.shape[0])
    m.reset(x0)
    out = m.solve_sens(x0, a, max_sqp=200)
    ok = golden["status"][:, 1] == 0
    assert np.array_equal(out["status"].cpu().numpy() == 0, ok)
    assert np.abs(out["u0"].cpu().numpy() - golden["a"]).max() == 0.0  # u0 is clamped to a
    assert _rel(out["cost"].cpu().numpy()[ok], golden["Q"][ok]) < 1e-9
    assert _rel(m.full_grad(out["dL"]).cpu().numpy()[ok], golden["dQ"][ok]) < 1e-6
    assert torch.count_nonzero(out["dpi"]) == 0  # dpi/dtheta == 0 in Q-mode by construction (quirk Q7)


def test_host_entry_point_equals_device_path(spec, golden):
    x0 = golden["x0"]
    m = _mpc(spec, x0.shape[0])
    m.reset(_dev(x0))
    o_dev = m.solve_sens(_dev(x0), max_sqp=100)
    m.reset(_dev(x0))
    o_host = m.solve_sens_host(x0, max_sqp=100)
    for k in ("u0", "cost", "dL", "dpi", "res"):
        assert np.array_equal(o_dev[k].cpu().numpy(), o_host[k]), k
    assert np.array_equal(o_dev["status"].cpu().numpy(), o_host["status"])
    # page-locked caller buffers are used as DMA source / target directly: same results
    m.reset(_dev(x0))
    pin = m.alloc_host_outputs(x0.shape[0], pinned=True)
    o_pin = m.solve_sens_host(torch.tensor(x0).pin_memory().numpy(), max_sqp=100, out=pin)
    for k in ("u0", "cost", "dL", "dpi", "res", "status"):
        assert np.array_equal(o_pin[k], o_host[k]), k


def test_live_oracle_kkt_and_sensitivities(spec):
    """The reference's own test method (nlp.py:1445-1537): the NLP mirror must accept the solver's
    primal-dual point as a KKT point, and the dense dR/dz solve must give the same dpi/dtheta."""
    from oracle.problems import make_cartpole
    from oracle.solver import DenseSolver

    rng = np.random.default_rng(7)
    x0 = rng.uniform([-1, -2, -np.pi, -4], [1, 2, np.pi, 4], size=(3, 4))
    m = _mpc(spec, 3)
    m.set_option("tol", 1e-10)
    m.reset(_dev(x0))
    out = m.solve_sens(_dev(x0), max_sqp=200)
    assert (out["status"].cpu().numpy() == 0).all()
    s = DenseSolver(make_cartpole("original"))
    N = spec.N
    for i in range(3):
        X = np.stack([m.get("x", k, 3)[i].cpu().numpy() for k in range(N + 1)])
        U = np.stack([m.get("u", k, 3)[i].cpu().numpy() for k in range(N)])
        sol, upd = s.unit(x0[i], init=(U, X), tol=1e-10)
        assert sol.sqp_iter <= 4  # the GPU point already is the oracle's KKT point (multipliers restart from 0)
        assert np.abs(sol.U - U).max() < 1e-7 and np.abs(sol.X - X).max() < 1e-7
        assert abs(sol.cost - out["cost"][i].item()) < 1e-9 * abs(sol.cost)
        assert _rel(m.full_grad(out["dL"][i]).cpu().numpy(), upd["dL_dp"][0]) < 1e-6
        assert _rel(m.full_grad(out["dpi"][i]).cpu().numpy(), upd["dpi_dp"]) < 1e-5


def test_rti_step_tracks_converged_solution(spec, golden):
    """K=1 (SQP-RTI) from a converged iterate at a nearby state: repeated RTI steps at the same state
    contract to the SQP solution (Gauss-Newton: linear rate), and an RTI step is exactly the first
    iteration of the SQP loop."""
    x0 = golden["x0"]
    B = x0.shape[0]
    ok = golden["status"][:, 0] == 0
    m = _mpc(spec, B)
    m.reset(_dev(x0))
    base = m.solve(_dev(x0), max_sqp=200)[0].cpu().numpy()
    x1 = x0 + 1e-4 * np.random.default_rng(0).standard_normal(x0.shape)
    errs = []
    us = []
    for _ in range(7):
        rti = m.solve_sens(_dev(x1), max_sqp=1)
        assert (rti["status"].cpu().numpy()[ok] == 0).all()
        us.append(rti["u0"].cpu().numpy())
    conv = m.solve_sens(_dev(x1), max_sqp=200)
    u_conv = conv["u0"].cpu().numpy()
    assert (conv["status"].cpu().numpy()[ok] == 0).all()
    errs = [np.abs(u - u_conv)[ok].max() for u in us]
    assert errs[0] < 1.0  # one step from a 1e-4 state change stays close (|u| <= 80)
    assert errs[3] < 0.05 * errs[0] and errs[6] < 1e-5  # linear contraction of the RTI iteration
    # the RTI step is the first SQP iteration: same start, max_sqp=1 vs the first of max_sqp=2
    m2 = _mpc(spec, B)
    m2.reset(_dev(x0))
    m2.solve(_dev(x0), max_sqp=200)
    a = m2.solve(_dev(x1), max_sqp=1)[0].cpu().numpy()
    assert np.abs(a - us[0])[ok].max() < 1e-12
    assert np.abs(base - golden["u0"])[ok].max() < 1e-6


def test_full_batch_properties(spec):
    """BASELINE size (65 536): every sample converges, KKT residuals tiny, replicated inputs give
    bit-identical outputs wherever they sit in the batch, dV/dtheta agrees with central differences
    of V over per-sample theta (the reference's FD check, scripts/linear_system_mpc_nlp.py:43-49)."""
    B = 65536
    g = torch.Generator(device="cpu").manual_seed(1234)
    lo = torch.tensor([-1.0, -2.0, -np.pi, -4.0], dtype=torch.float64)
    x0 = (lo + (-2 * lo) * torch.rand(B, 4, generator=g, dtype=torch.float64)).cuda()
    x0[B // 2:] = x0[: B // 2]  # replicate
    m = _mpc(spec, B)
    m.set_option("tol", 1e-8)
    m.reset(x0)
    out = m.solve_sens(x0, max_sqp=200)
    st = out["status"].cpu().numpy()
    assert (st == 0).mean() > 0.9, np.bincount(st)  # full-step SQP (acados default) 2-cycles on a few % of random states
    okm = out["status"] == 0
    assert out["res"][okm].max().item() < 1e-8 * (1 + 1e-3)  # converged at tol=1e-8; re-evaluated by sens
    for k in ("u0", "cost", "dL", "dpi"):
        assert torch.equal(out[k][: B // 2], out[k][B // 2:]), k
    # FD check of dV/dtheta through per-sample theta
    n = 64
    xs = x0[:n]
    d = 1e-6
    th = np.tile(spec.p_nominal, (6 * n, 1))
    for j in range(3):
        th[(2 * j) * n:(2 * j + 1) * n, j] += d
        th[(2 * j + 1) * n:(2 * j + 2) * n, j] -= d
    m2 = _mpc(spec, 6 * n)
    m2.set_option("tol", 1e-11)
    m2.set_theta(th)
    xx = xs.repeat(6, 1)
    m2.reset(xx)
    o2 = m2.solve_sens(xx, max_sqp=300)
    V = o2["cost"].cpu().numpy().reshape(6, n)
    good = (o2["status"].cpu().numpy().reshape(6, n) == 0).all(axis=0) & (st[:n] == 0)
    assert good.sum() > n // 2
    fd = np.stack([(V[2 * j] - V[2 * j + 1]) / (2 * d) for j in range(3)], axis=1)
    an = out["dL"][:n, :3].cpu().numpy()
    assert np.abs(fd - an)[good].max() < 1e-4 * max(1.0, np.abs(an[good]).max())


def test_td_grad_reduction(spec):
    B = 1000
    m = _mpc(spec, B)
    g = torch.Generator(device="cpu").manual_seed(3)
    td = torch.randn(B, generator=g, dtype=torch.float64).cuda()
    dQ = torch.randn(B, spec.ntheta, generator=g, dtype=torch.float64).cuda()
    status = (torch.rand(B, generator=g) < 0.1).to(torch.int32).cuda() * 2
    acc = m.td_grad(td, dQ, status)
    ok = status == 0
    ref = torch.cat([(td[ok, None] * dQ[ok]).sum(0), td[ok].sum()[None], ok.sum()[None].double()])
    assert torch.allclose(acc, ref, rtol=1e-12, atol=1e-10)


def test_errors_are_reported(spec):
    from mpc4rl_b200 import BatchedMPC

    m = BatchedMPC(spec, max_batch=4, device=0)
    with pytest.raises(RuntimeError):
        m.solve(torch.zeros(8, 4, dtype=torch.float64, device="cuda:0"))  # batch > max_batch
    with pytest.raises(TypeError):
        m.solve(torch.zeros(2, 4, dtype=torch.float32, device="cuda:0"))
    with pytest.raises(RuntimeError):
        m.set_option("no_such_option", 1.0)


def test_warp_per_sample_queue_equals_thread_per_sample(spec):
    """The warp-cooperative queue kernel (csrc/coop.cuh, option coop=1, the default for this model) runs
    the interior-point iteration of the thread-per-sample kernels statement by statement: same statuses,
    u0 / V / gradients equal to rounding -- from a cold start (SQP to convergence: every sample goes
    through the queue in the first rounds), for RTI steps at perturbed states, and in Q-mode."""
    B = 4096
    g = torch.Generator(device="cpu").manual_seed(7)
    lo = torch.tensor([-1.0, -2.0, -np.pi, -4.0], dtype=torch.float64)
    x0 = (lo + (-2 * lo) * torch.rand(B, 4, generator=g, dtype=torch.float64)).cuda()
    dx = 1e-3 * torch.randn(3, B, 4, generator=g, dtype=torch.float64).cuda()
    a0 = (-80.0 + 160.0 * torch.rand(B, 1, generator=g, dtype=torch.float64)).cuda()
    runs = {}
    for coop in (1, 0):
        m = _mpc(spec, B)
        m.set_option("tol", 1e-8)
        m.set_option("coop", coop)
        m.reset(x0)
        n0 = m.launch_count
        outs = [m.solve_sens(x0, max_sqp=60)]
        for i in range(3):
            outs.append(m.solve_sens(x0 + dx[i], max_sqp=1))
        outs.append(m.solve_sens(x0, u0=a0, max_sqp=60))
        runs[coop] = [{k: v.cpu().numpy() for k, v in o.items()} for o in outs]
        runs[coop].append(m.launch_count - n0)
    assert runs[1][-1] < runs[0][-1]  # one queue kernel instead of gather + condense + solve + scatter
    conv = None  # samples converged in the cold solve: only those start the later calls from the same iterate
    for step, (a, b) in enumerate(zip(runs[1][:-1], runs[0][:-1])):
        same = a["status"] == b["status"]
        assert same.mean() > 0.995, (step, same.mean())
        ok = same & (a["status"] == 0)
        conv = ok if conv is None else conv
        ok = ok & conv
        assert ok.mean() > 0.85, (step, ok.mean())
        assert np.abs(a["u0"] - b["u0"])[ok].max() < 1e-7, step
        assert _rel(a["cost"][ok], b["cost"][ok]) < 1e-9, step
        assert _rel(a["dL"][ok], b["dL"][ok]) < 1e-6, step
        if step < 4:  # Q-mode: dpi/dtheta is ~0 by construction (quirk Q7)
            assert _rel(a["dpi"][ok], b["dpi"][ok]) < 1e-5, step


def test_phase_timings_and_queue_statistics(spec):
    """rlmpc_get_timings: six phase times of the last call plus the queue length and the interior-point
    iterations spent on it (the reference's analogue is update_nlp's nlp_timing dict, nlp.py:1397-1422)."""
    B = 2048
    g = torch.Generator(device="cpu").manual_seed(3)
    lo = torch.tensor([-1.0, -2.0, -np.pi, -4.0], dtype=torch.float64)
    x0 = (lo + (-2 * lo) * torch.rand(B, 4, generator=g, dtype=torch.float64)).cuda()
    m = _mpc(spec, B)
    with pytest.raises(RuntimeError):
        m.timings()  # option "timing" is off
    m.set_option("timing", 1)
    m.reset(x0)
    m.solve(x0, max_sqp=60)
    m.solve_sens(x0 + 1e-3, max_sqp=1)
    t = m.timings()
    assert set(m.PHASES) <= set(t) and all(t[k] >= 0.0 for k in m.PHASES)
    assert t["linearize"] > 0.0 and t["qp_fast"] > 0.0 and t["sens_sweep"] > 0.0
    assert 0 < t["queue_len"] < B  # some, not all, samples need more than the one warm Newton iteration
    assert t["queue_ipm_iters"] >= 2 * t["queue_len"]  # each of them at least a second iteration


def test_handles_with_different_horizons_coexist(spec):
    """Kernel attributes (dynamic shared memory of the warp-per-sample queue kernel) are per kernel, not
    per handle: a long-horizon and a short-horizon handle alive at the same time, used alternately,
    give what each gives alone."""
    import copy

    from mpc4rl_b200 import BatchedMPC, cartpole_original_config, cartpole_spec

    B = 512
    g = torch.Generator(device="cpu").manual_seed(11)
    lo = torch.tensor([-1.0, -2.0, -np.pi, -4.0], dtype=torch.float64)
    x0 = (lo + (-2 * lo) * torch.rand(B, 4, generator=g, dtype=torch.float64)).cuda()
    cfgs = []
    for N in (80, 20):
        c = copy.deepcopy(cartpole_original_config())
        c["dimensions"]["N"] = N
        c["ocp_options"]["tf"] = 0.02 * N
        cfgs.append(cartpole_spec(c))

    def run(m):
        m.reset(x0)
        o = m.solve_sens(x0, max_sqp=40)
        return {k: v.clone() for k, v in o.items()}

    alone = []
    for s in cfgs:
        m = BatchedMPC(s, max_batch=B, device=0)
        alone.append(run(m))
        m.close()
    both = [BatchedMPC(s, max_batch=B, device=0) for s in cfgs]  # the short horizon is created last
    for rep in range(2):
        for m, ref in zip(both, alone):
            o = run(m)
            assert (o["status"] == 0).double().mean().item() > 0.5  # N=80 from random states: a third hits the 40-iteration limit
            for k in ("status", "u0", "cost", "dL", "dpi"):
                assert torch.equal(o[k], ref[k]), k


def test_split_batch_on_two_streams_is_bit_identical(spec):
    """Option split (default 2): an RTI call runs the two halves of the batch as two chains of kernels on two
    streams.  Samples are independent, so every output is bit-identical to the single-chain call, also for
    a batch that does not divide into whole tiles."""
    B = 8192 + 40
    g = torch.Generator(device="cpu").manual_seed(21)
    lo = torch.tensor([-1.0, -2.0, -np.pi, -4.0], dtype=torch.float64)
    x0 = (lo + (-2 * lo) * torch.rand(B, 4, generator=g, dtype=torch.float64)).cuda()
    dx = 1e-3 * torch.randn(2, B, 4, generator=g, dtype=torch.float64).cuda()
    a0 = (-80.0 + 160.0 * torch.rand(B, 1, generator=g, dtype=torch.float64)).cuda()
    runs = []
    for split in (2, 1, 3):
        m = _mpc(spec, B)
        m.set_option("tol", 1e-8)
        m.set_option("split", split)
        m.set_option("timing", 1)
        m.reset(x0)
        m.solve(x0, max_sqp=40)
        outs = [m.solve_sens(x0 + dx[0], max_sqp=1), m.solve_sens(x0 + dx[1], u0=a0, max_sqp=1)]
        q = m.timings()["queue_len"]
        outs.append(dict(zip(("u0", "cost", "status"), m.solve(x0 + dx[0], max_sqp=1))))
        runs.append(([{k: v.clone() for k, v in o.items()} for o in outs], q))
    assert runs[0][1] == runs[1][1] == runs[2][1] > 0  # queue statistics add up over the parts
    for other in (runs[1], runs[2]):
        for a, b in zip(runs[0][0], other[0]):
            for k in a:
                assert torch.equal(a[k], b[k]), k


def test_pipelined_host_entry_point_is_bit_identical(spec):
    """rlmpc_solve_sens_host with page-locked buffers runs an RTI call as per-part pipelines (copy in, kernel
    chain, copy out) on separate streams; same numbers as the device-resident call, V- and Q-mode, ragged batch."""
    B = 8192 + 72
    g = torch.Generator(device="cpu").manual_seed(22)
    lo = torch.tensor([-1.0, -2.0, -np.pi, -4.0], dtype=torch.float64)
    x0 = lo + (-2 * lo) * torch.rand(B, 4, generator=g, dtype=torch.float64)
    x1 = (x0 + 1e-3 * torch.randn(B, 4, generator=g, dtype=torch.float64)).pin_memory()
    a0 = (-80.0 + 160.0 * torch.rand(B, 1, generator=g, dtype=torch.float64)).pin_memory()
    res = []
    for host in (False, True):
        m = _mpc(spec, B)
        m.set_option("tol", 1e-8)
        m.reset(x0.cuda())
        m.solve(x0.cuda(), max_sqp=40)
        outs = []
        for u in (None, a0):
            if host:
                pin = m.alloc_host_outputs(B, pinned=True)
                o = m.solve_sens_host(x1.numpy(), None if u is None else u.numpy(), max_sqp=1, out=pin)
                outs.append({k: np.array(v) for k, v in o.items()})
            else:
                o = m.solve_sens(x1.cuda(), None if u is None else u.cuda(), max_sqp=1)
                outs.append({k: v.cpu().numpy() for k, v in o.items()})
        res.append(outs)
    for a, b in zip(res[0], res[1]):
        for k in ("u0", "cost", "dL", "dpi", "res", "status"):
            assert np.array_equal(a[k], b[k]), k


def test_cost_parameter_columns_match_oracle(spec):
    """parameterize_tracking_cost=True: W_0, W, W_e, yref_0, yref, yref_e carry gradient -- all 83 columns of dL/dp AND
    of dpi/dp (the latter contracted on the fly in the adjoint sweep) against tests/golden/cartpole_paramcost.npz."""
    from mpc4rl_b200 import BatchedMPC

    g = np.load(os.path.join(ROOT, "tests", "golden", "cartpole_paramcost.npz"))
    B = g["x0"].shape[0]
    m = BatchedMPC(spec, max_batch=B, device=0)
    m.set_option("tol", 1e-10)
    m.set_option("param_cost", 1)
    m.set_theta(g["theta"])
    x0 = _dev(g["x0"])
    m.reset(x0)
    out = m.solve_sens(x0, max_sqp=200)
    ok = (out["status"].cpu().numpy() == 0) & (g["status"][:, 0] == 0)
    assert ok.sum() >= 4
    assert out["dL"].shape[1] == 83 and out["dpi"].shape[2] == 83
    assert np.abs(out["u0"].cpu().numpy() - g["u0"])[ok].max() < 1e-6
    assert _rel(out["dL"].cpu().numpy()[ok], g["dV"][ok]) < 1e-6
    assert np.abs(g["dpi"][ok][:, :, 3:]).max() > 1.0
    assert _rel(out["dpi"].cpu().numpy()[ok], g["dpi"][ok]) < 1e-5
    # a second call into the same output buffers must not accumulate on top of the first
    out2 = m.solve_sens(x0, max_sqp=1, out=out)
    assert _rel(out2["dpi"].cpu().numpy()[ok], g["dpi"][ok]) < 1e-5


def test_free_g_parameter_set_matches_oracle():
    """g un-fixed as scripts/cartpole_mpc_qlearning.py:184-187 does: 84 parameters, 4 with gradient, against
    tests/golden/cartpole_free_g.npz (V- and Q-mode)."""
    import copy

    from mpc4rl_b200 import BatchedMPC, cartpole_original_config, cartpole_spec

    g = np.load(os.path.join(ROOT, "tests", "golden", "cartpole_free_g.npz"))
    cfg = copy.deepcopy(cartpole_original_config())
    cfg["model"]["params"]["g"]["fixed"] = False
    fspec = cartpole_spec(cfg)
    assert fspec.ntheta == 84
    B = g["x0"].shape[0]
    m = BatchedMPC(fspec, max_batch=B, device=0)
    m.set_option("tol", 1e-10)
    x0 = _dev(g["x0"])
    m.reset(x0)
    out = m.solve_sens(x0, max_sqp=200)
    ok = (out["status"].cpu().numpy() == 0) & (g["status"][:, 0] == 0)
    assert ok.sum() >= 4
    assert out["dL"].shape[1] == 4 and np.abs(g["dV"][ok][:, 3]).max() > 1e-3
    assert np.abs(out["u0"].cpu().numpy() - g["u0"])[ok].max() < 1e-6
    assert _rel(out["cost"].cpu().numpy()[ok], g["V"][ok]) < 1e-9
    assert _rel(m.full_grad(out["dL"]).cpu().numpy()[ok], g["dV"][ok]) < 1e-6
    assert _rel(m.full_grad(out["dpi"]).cpu().numpy()[ok], g["dpi"][ok]) < 1e-5
    m.reset(x0)
    q = m.solve_sens(x0, u0=_dev(g["a"]), max_sqp=200)
    okq = (q["status"].cpu().numpy() == 0) & (g["status"][:, 1] == 0)
    assert okq.sum() >= 4
    assert _rel(q["cost"].cpu().numpy()[okq], g["Q"][okq]) < 1e-9
    assert _rel(m.full_grad(q["dL"]).cpu().numpy()[okq], g["dQ"][okq]) < 1e-6
    # a changed g moves the solution: the parameter is live in the model, not only in the gradient
    th = g["theta"].copy()
    th[3] = 5.0
    m.set_theta(th)
    m.reset(x0)
    o2 = m.solve_sens(x0, max_sqp=200)
    assert np.abs(o2["cost"].cpu().numpy() - g["V"])[ok].max() > 1e-3


@pytest.mark.parametrize("B", [512, 8192])
def test_cuda_graph_replay_is_bit_identical(spec, B):
    """Option "graph": the RTI kernel chain (two streams for B >= 4096) captured once and replayed gives exactly the
    results of the eagerly launched chain, step after step, also when the inputs change in place."""
    from mpc4rl_b200 import BatchedMPC

    torch.manual_seed(3)
    lo = torch.tensor([-1.0, -2.0, -np.pi, -4.0], dtype=torch.float64)
    x0 = (lo + (-2.0 * lo) * torch.rand(B, 4, dtype=torch.float64)).cuda()
    outs = []
    for graph in (0, 1):
        m = BatchedMPC(spec, max_batch=B, device=0)
        m.set_option("graph", graph)
        m.reset(x0)
        m.solve(x0, max_sqp=50)
        x = x0.clone()
        out = m.alloc_outputs(B)
        res = []
        for i in range(4):
            x.add_(1e-2 * torch.sin(torch.arange(B * 4, device="cuda", dtype=torch.float64).reshape(B, 4) + i))
            m.solve_sens(x, max_sqp=1, out=out)
            res.append({k: v.clone() for k, v in out.items()})
        outs.append(res)
    for a, b in zip(*outs):
        for k in ("u0", "cost", "dL", "dpi", "status", "res"):
            assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("opt", [("ipm_passes", 3), ("fuse_lin", 1)])
def test_opt_in_kernel_variants_give_the_same_results(spec, opt):
    """The measured-and-rejected variants stay selectable (profiles/r02_summary.md): pass kernels ahead of the queue kernel
    (option ipm_passes) and the linearisation inside the fast-path kernel (option fuse_lin).  Same statuses, outputs equal
    to rounding, over a cold SQP solve and closed-loop-sized RTI steps."""
    B = 4096
    g = torch.Generator(device="cpu").manual_seed(21)
    lo = torch.tensor([-1.0, -2.0, -np.pi, -4.0], dtype=torch.float64)
    x0 = (lo + (-2 * lo) * torch.rand(B, 4, generator=g, dtype=torch.float64)).cuda()
    dx = 2e-2 * torch.randn(2, B, 4, generator=g, dtype=torch.float64).cuda()
    runs = []
    for on in (False, True):
        m = _mpc(spec, B)
        m.set_option("tol", 1e-8)
        if on:
            m.set_option(opt[0], opt[1])
        m.reset(x0)
        outs = [m.solve_sens(x0, max_sqp=60)] + [m.solve_sens(x0 + dx[i], max_sqp=1) for i in range(2)]
        runs.append([{k: v.cpu().numpy() for k, v in o.items()} for o in outs])
    conv = None
    for step, (a, b) in enumerate(zip(*runs)):
        same = a["status"] == b["status"]
        assert same.mean() > 0.995, (step, same.mean())
        ok = same & (a["status"] == 0)
        conv = ok if conv is None else conv
        ok = ok & conv
        assert ok.mean() > 0.8, (step, ok.mean())
        assert np.abs(a["u0"] - b["u0"])[ok].max() < 1e-7, step
        assert _rel(a["cost"][ok], b["cost"][ok]) < 1e-9, step
        assert _rel(a["dpi"][ok], b["dpi"][ok]) < 1e-5, step
