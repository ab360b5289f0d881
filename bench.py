#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric   MPC solves+sensitivities/sec: one unit = for one sample, one SQP-RTI step of the NMPC from the stored
         warm-start iterate PLUS dL/dtheta and dpi/dtheta at the new iterate (= one `update`/`q_update` + one
         `update_nlp` of the reference).  SURVEY.md 8(d).
step     one rlmpc_solve_sens call over one batch of synthetic samples per GPU, followed by the TD-gradient accumulator
         kernel and, for N > 1, its NCCL all-reduce.  Headline workload (`--workload cartpole`, BASELINE.json
         configs[1]): 65 536 cart-pole states per GPU that move by ONE ENVIRONMENT STEP between the timed steps
         (x+ = env.step(x, pi(x)), rlmpc_cartpole_env_step, tau = 0.02 -- SURVEY.md 8(d) "converged iterate perturbed
         by one env step").  Other workloads: see WORKLOADS.
value    whole-job units/s with inputs resident in HBM, CUDA-event timed, max over ranks.
e2e      same metric through the C ABI host entry point (rlmpc_solve_sens_host): pinned host buffers, H2D of the states
         and D2H of every result inside the timed region (wall clock, max over ranks), replaying the same states.
roofline HBM fraction as defined by SURVEY.md 8(d) (algorithmic bytes) AND the FP64 fraction: executed FP64 flops per
         unit (ncu instruction counts, profiles/) over the FMA throughput measured on this device by rlmpc_fp64_peak.

`--impl reference` times the CPU path instead: acados/CasADi are not installable here, so this is the oracle's host
port of the same structure-exploiting algorithm (oracle/cpu_port, all host threads) -- "restated, not acados".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MPC solves+sensitivities/sec (cartpole N=40, batch 65k)"
UNIT = "units/s"
ENV = dict(gravity=9.8, masscart=1.0, masspole=0.1, length=0.5, force_mag=80.0, tau=0.02)


def synth_states(B, seed):
    """s~U(-1,1), s_dot~U(-2,2), theta~U(-pi,pi), theta_dot~U(-4,4)."""
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    lo = torch.tensor([-1.0, -2.0, -np.pi, -4.0], dtype=torch.float64)
    return lo + (-2.0 * lo) * torch.rand(B, 4, generator=g, dtype=torch.float64)


def env_step_np(x, force):
    """rlmpc/gym/continuous_cartpole/environment.py:104-131 (euler), vectorised; force in newtons (CPU arm)."""
    g, mc, mp_, ln, tau = ENV["gravity"], ENV["masscart"], ENV["masspole"], ENV["length"], ENV["tau"]
    total, pml = mc + mp_, mp_ * ln
    s, sd, th, thd = x.T
    f = np.clip(force, -ENV["force_mag"], ENV["force_mag"])
    ct, sn = np.cos(th), np.sin(th)
    temp = (f + pml * thd**2 * sn) / total
    thacc = (g * sn - ct * temp) / (ln * (4.0 / 3.0 - mp_ * ct**2 / total))
    xacc = temp - pml * thacc * ct / total
    return np.stack([s + tau * sd, sd + tau * xacc, th + tau * thd, thd + tau * thacc], 1)


_CP = ("cartpole_original N=40 nx=4 nu=1 ntheta=83 (3 with gradient), V-mode SQP-RTI (K=1) + dL/dtheta + dpi/dtheta, "
       "warm-started from the previous iterate")
WORKLOADS = {
    # BASELINE.json configs[1] -- the headline.  b_alg: SURVEY.md 8(d), 8*(2*524 + 4+1 + 1+1+3+3) + 4
    "cartpole": dict(batch=65536, b_alg=8492, metric=METRIC,
                     text=_CP + "; closed loop: the states move by one environment step under the MPC policy between steps "
                                "(rlmpc_cartpole_env_step, tau=0.02, force_mag=80, no episode resets)"),
    # the round-1 workload, kept for continuity: converged states + 1e-3 * randn, a 40-80 x smaller move
    "cartpole_tiny_pert": dict(batch=65536, b_alg=8492, metric=METRIC + " [tiny perturbation variant]",
                               text=_CP + "; states = converged states + 1e-3 * randn every step (round-1 workload)"),
    # replay-buffer shape: every step draws a random minibatch of a 4x larger iterate store (rlmpc_store_copy inside the
    # timed region) and evaluates it one environment step away from the state its iterate was converged at
    "cartpole_replay": dict(batch=65536, b_alg=8492, metric=METRIC + " [replay-buffer variant]",
                            text=_CP + "; every step gathers a random minibatch of 65536 iterates out of a store of 262144 "
                                       "(rlmpc_store_copy, timed) and solves it at the stored state moved by one environment step"),
    # SURVEY.md 8(d) config 2, secondary: config/cartpole.yaml (N=30, box bounds on all states)
    "cartpole_bx": dict(batch=65536, b_alg=14092, metric="MPC solves+sensitivities/sec (cartpole.yaml N=30 with state bounds, batch 65k)",
                        text="cartpole.yaml N=30 nx=4 nu=1, input and state bounds, V-mode SQP-RTI (K=1) + dL/dtheta + dpi/dtheta; "
                             "feasible starts (closed-loop rollout states), states perturbed by 1e-3 * randn every step"),
    # BASELINE.json configs[3]: N=100, nu=3 (third input = slack), affine h rows
    "evaporation": dict(batch=32768, b_alg=45228, metric="MPC solves+sensitivities/sec (evaporation N=100, batch 32k)",
                        text="evaporation_process N=100 nx=2 nu=3 ntheta=60 (tracking-cost parameters), V-mode SQP-RTI "
                             "(K=1) + dL/dtheta + dpi/dtheta, warm-started from the converged iterate, states perturbed every step"),
    # BASELINE.json configs[2]: chain of masses, reference default n_mass = 5 (nx = 21) and BASELINE's nx = 27 (n_mass = 6)
    "chain_mass": dict(batch=8192, b_alg=53012, n_mass=5, metric="MPC solves+sensitivities/sec (chain_mass n_mass=5 nx=21 N=40, batch 8192)",
                       text="chain_mass n_mass=5 nx=21 nu=3 N=40 ntheta=499 (all with gradient), V-mode SQP-RTI (K=1) + dL/dtheta + "
                            "dpi/dtheta, warm-started from the previous iterate; x0 = define_x0 + N(0, 1e-2) redrawn every step"),
    "chain_mass_6": dict(batch=8192, b_alg=70468, n_mass=6, metric="MPC solves+sensitivities/sec (chain_mass n_mass=6 nx=27 N=40, batch 8192)",
                         text="chain_mass n_mass=6 nx=27 nu=3 N=40 ntheta=800 (all with gradient), V-mode SQP-RTI (K=1) + dL/dtheta + "
                              "dpi/dtheta, warm-started from the previous iterate; x0 = define_x0 + N(0, 1e-2) redrawn every step"),
}
# Executed FP64 flops per unit: 2 * DFMA + DADD + DMUL thread-level instructions summed over the kernels of one step and
# divided by the units of the step -- ncu counters of the committed step profiles (profiles/r02z_step_<workload>.csv (final build of round 2),
# tools/profile_step.py, summarised in profiles/r02_summary.md).  The chain-mass kernels also run FP64 tensor-core
# instructions (mma.sync m8n8k4 = 512 flop per warp instruction, zero padding of 21 -> 24 included); their count is
# structural: per sample 40 stages x 144 per Riccati factorisation (one per interior-point iteration of the step + one
# in the sensitivity sweep) and 40 x 8 x 27 in the Hessian accumulation.  None = not measured for this workload.
FLOPS_PER_UNIT = {"cartpole": 373203.0, "cartpole_tiny_pert": 236867.0, "evaporation": 1161236.0, "chain_mass": 10111475.0}
FLOPS_SOURCE = "profiles/r02z_step_*.csv (smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum per kernel) + structural DMMA count"
# dram__bytes_read.sum + dram__bytes_write.sum of the kernels of one step (same profiles)
TRAFFIC_PER_STEP = {"cartpole": 8.430e9, "cartpole_tiny_pert": 8.230e9, "evaporation": 11.260e9, "chain_mass": 29.114e9}


def dmma_flops_per_unit(workload, ipm_iters_mean):
    if workload != "chain_mass":
        return 0.0
    return 512.0 * (40 * 144 * (ipm_iters_mean + 1.0) + 40 * 8 * 27)


def config_dict(B, n_gpus, workload="cartpole", scaling="weak", graph=False):
    return {"workload": WORKLOADS[workload]["text"],
            "batch_per_gpu": B, "global_batch": B * n_gpus, "parallelism": f"dp{n_gpus} (batch shards, replicated theta)",
            "scaling": scaling, "cuda_graph": bool(graph),
            "l2": "working set (iterate + stage scratch, > 1.5 GB per GPU) is larger than L2, no flush needed",
            "seed": 1234}


class Workload:
    """Problem + the sequence of states the timed steps see.  `advance(i, out)` returns the state tensor of step i (and
    may launch kernels: the environment step of the closed loop, the gather from the iterate store)."""

    def __init__(self, name, B, rank, dev):
        import torch

        self.name, self.B, self.dev, self.rank = name, B, dev, rank
        self.store = None
        self.pre = None
        if name.startswith("cartpole") and name != "cartpole_bx":
            from mpc4rl_b200 import cartpole_original_config, cartpole_spec

            self.spec = cartpole_spec(cartpole_original_config())
            self.x0 = synth_states(B, 1234 + rank).to(dev)
            self.pert = 1e-3
        elif name == "cartpole_bx":
            from mpc4rl_b200 import cartpole_config, cartpole_spec

            self.spec = cartpole_spec(cartpole_config())
            self.x0 = None  # feasible starts are produced by a closed-loop rollout in setup()
            self.pert = 1e-3
        elif name == "evaporation":
            from mpc4rl_b200 import evaporation_spec

            self.spec = evaporation_spec(gamma=0.99)
            g = torch.Generator(device="cpu").manual_seed(7 + rank)
            lo, hi = torch.tensor([25.0, 49.7], dtype=torch.float64), torch.tensor([40.0, 70.0], dtype=torch.float64)
            self.x0 = (lo + (hi - lo) * torch.rand(B, 2, generator=g, dtype=torch.float64)).to(dev)
            self.pert = 1e-2
        else:
            from mpc4rl_b200.problems import chain_define_x0, chain_mass_spec, get_chain_params

            cp = get_chain_params()
            cp["n_mass"] = WORKLOADS[name]["n_mass"]
            self.spec = chain_mass_spec(cp)
            g = torch.Generator(device="cpu").manual_seed(50 + rank)  # seed 50, perturb_scale 1e-2 (ocp_utils.py:334-339)
            self.xbase = torch.tensor(chain_define_x0(cp), dtype=torch.float64)
            self.x0 = (self.xbase + 1e-2 * torch.randn(B, self.spec.nx, generator=g, dtype=torch.float64)).to(dev)
            self.pert = 1e-2

    # ---- untimed setup: the warm-start state every timed run starts from ----
    def install_guess(self, mpc):
        import torch

        if self.name == "evaporation":  # every stage on the steady state (evaporation_process/acados.py:104-109)
            B, spec, dev = self.B, self.spec, self.dev
            mpc.reset(B=B)
            xs = torch.tensor(spec.x_init, dtype=torch.float64, device=dev).repeat(B, 1)
            us = torch.tensor(spec.u_init, dtype=torch.float64, device=dev).repeat(B, 1)
            for k in range(spec.N + 1):
                mpc.put("x", k, xs)
            for k in range(spec.N):
                mpc.put("u", k, us)
        else:
            mpc.reset(self.x0)

    def setup(self, mpc, n_steps):
        """Converge the batch once (cold start like MPC.reset); returns the fraction of converged samples."""
        import torch

        B, dev = self.B, self.dev
        if self.name == "cartpole_bx":
            # feasible starts: roll the hanging pole (config/cartpole.yaml x0 = [0,0,3.14,0] + noise) out in closed loop for a
            # random number of steps under the MPC policy; every visited state is inside the state box
            g = torch.Generator(device="cpu").manual_seed(1234 + self.rank)
            x = torch.tensor([0.0, 0.0, np.pi, 0.0], dtype=torch.float64) + torch.tensor([0.5, 0.5, 0.3, 0.5], dtype=torch.float64) * (
                2.0 * torch.rand(B, 4, generator=g, dtype=torch.float64) - 1.0)
            self.x0 = x.to(dev)
            mpc.reset(self.x0)
            mpc.solve(self.x0, max_sqp=60)
            stop = torch.randint(0, 25, (B,), generator=g).to(dev)
            env = self._env(force_mag=30.0)
            xc = self.x0.clone()
            for i in range(25):
                u0, _, _ = mpc.solve(xc, max_sqp=2)
                xn = xc.clone()
                env(xn, (u0[:, 0] / 30.0).clamp(-1.0, 1.0))
                move = (stop > i).unsqueeze(1)
                xc = torch.where(move, xn, xc)
            self.x0 = xc
        self.install_guess(mpc)
        _, _, st = mpc.solve(self.x0, max_sqp=60)
        torch.cuda.synchronize()
        g = torch.Generator(device="cpu").manual_seed(99 + self.rank)
        if self.name in ("cartpole", "cartpole_replay"):
            self.env = self._env()
            self.xcur = self.x0.clone()
            self.log = torch.empty(n_steps, B, self.spec.nx, dtype=torch.float64, device=dev)
        if self.name == "cartpole_replay":
            cap = 4 * B
            self.store = mpc.iterate_store(cap)
            self.store_x1 = torch.empty(cap, self.spec.nx, dtype=torch.float64, device=dev)
            for c in range(4):  # fill the store: converged iterates of 4 B states, and each state moved by one env step
                xs = synth_states(B, 4321 + 10 * self.rank + c).to(dev)
                mpc.reset(xs)
                u0, _, _ = mpc.solve(xs, max_sqp=60)
                idx = torch.arange(c * B, (c + 1) * B, device=dev, dtype=torch.int32)
                self.store.save(idx)
                x1 = xs.clone()
                self.env(x1, (u0[:, 0] / ENV["force_mag"]).clamp(-1.0, 1.0))
                self.store_x1[c * B:(c + 1) * B] = x1
            self.idx = [torch.randperm(cap, generator=g)[:B].to(dev, torch.int32) for _ in range(n_steps)]
            self.idx_long = [i.long() for i in self.idx]
            self.xbuf = torch.empty(B, self.spec.nx, dtype=torch.float64, device=dev)
            self.install_guess(mpc)
            mpc.solve(self.x0, max_sqp=60)
        if self.name in ("cartpole_tiny_pert", "cartpole_bx", "evaporation"):
            self.pre = [(self.x0 + self.pert * torch.randn(B, self.spec.nx, generator=g, dtype=torch.float64).to(dev)) for _ in range(n_steps)]
        if self.name.startswith("chain_mass"):
            self.pre = [(self.xbase + 1e-2 * torch.randn(B, self.spec.nx, generator=g, dtype=torch.float64)).to(dev) for _ in range(n_steps)]
        torch.cuda.synchronize()
        return float((st == 0).double().mean().item())

    def _env(self, force_mag=ENV["force_mag"]):
        """state[B,4] <- env.step(state, action[B] in [-1,1]) on the device, no terminations / resets."""
        import ctypes as C

        import torch

        from mpc4rl_b200 import _cabi

        lib, B, dev = _cabi.load(), self.B, self.dev
        par = torch.tensor([ENV["gravity"], ENV["masscart"], ENV["masspole"], ENV["length"], force_mag, ENV["tau"], 1e30, 1e30, 2e9,
                            0.0, 0.0, np.pi, 0.0], dtype=torch.float64, device=dev)
        rew = torch.empty(B, dtype=torch.float64, device=dev)
        term, trunc, steps = (torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(3))
        p = lambda t: C.c_void_p(t.data_ptr())

        def step(state, action):
            a = action.contiguous()
            _cabi.check(lib.rlmpc_cartpole_env_step(p(par), B, p(state), p(a), p(rew), p(term), p(trunc), p(steps),
                                                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            self._keep = (par, rew, term, trunc, steps, a)

        return step

    def rewind(self, mpc):
        """Back to the state after setup (for the replays: per-kernel pass, e2e pass)."""
        self.install_guess(mpc)
        mpc.solve(self.x0, max_sqp=60)
        if self.name in ("cartpole", "cartpole_replay"):
            self.xcur = self.x0.clone()

    def advance(self, i, out, mpc, record=True):
        import torch

        if self.name == "cartpole":
            if i > 0:  # the state moves under the policy of the previous step
                self.env(self.xcur, (out["u0"][:, 0] / ENV["force_mag"]).clamp(-1.0, 1.0))
            if record:
                self.log[i].copy_(self.xcur)
            return self.xcur
        if self.name == "cartpole_replay":
            self.store.load(self.idx[i])
            # (no allocation inside the timed region: preallocated index / state buffers)
            x = torch.index_select(self.store_x1, 0, self.idx_long[i], out=self.xbuf)
            if record:
                self.log[i].copy_(x)
            return x
        return self.pre[i]

    def logged(self, i):
        return self.log[i] if self.pre is None else self.pre[i]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_port_run(workload, n_samples, steps, warmup, threads=0, seed=1234):
    """RTI + sensitivities on the host (oracle/cpu_port): returns (units/s, threads, seconds per step, description)."""
    from oracle import cpu_port as cp

    if workload.startswith("chain_mass"):
        from mpc4rl_b200.problems import chain_define_x0, chain_mass_spec, get_chain_params

        cpar = get_chain_params()
        cpar["n_mass"] = WORKLOADS[workload]["n_mass"]
        spec = chain_mass_spec(cpar)
        pd = cp.make_pd(spec.N, spec.cost_scaling(), spec.lbu, spec.ubu, spec.model_const, tol=1e-6, warm_ipm=1)
        rng = np.random.default_rng(seed)
        xb = chain_define_x0(cpar)
        x0 = xb + 1e-2 * rng.standard_normal((n_samples, spec.nx))
        o = cp.chain_unit(cpar["n_mass"], pd, 0, 50, spec.p_nominal, spec.x_ss, x0, do_sens=False, threads=threads)
        it, ts = o["iterate"], []
        for s in range(warmup + steps):
            x1 = xb + 1e-2 * rng.standard_normal((n_samples, spec.nx))
            t0 = time.perf_counter()
            o = cp.chain_unit(cpar["n_mass"], pd, 0, 1, spec.p_nominal, spec.x_ss, x1, iterate=it, threads=threads)
            dt = time.perf_counter() - t0
            it = o["iterate"]
            if s >= warmup:
                ts.append(dt)
        nthreads = threads if threads else (os.cpu_count() or 1)
        return (n_samples * len(ts) / sum(ts), min(nthreads, n_samples), sum(ts) / len(ts),
                "host run of the warp-cooperative engine under a fiber emulation of a warp (oracle/cpu_port/chain_port.cpp): "
                "same algorithm, NOT acados and not a tuned CPU code")
    from mpc4rl_b200.problems import cartpole_original_config, cartpole_spec

    spec = cartpole_spec(cartpole_original_config())
    pd = cp.make_pd(spec.N, spec.cost_scaling(), spec.lbu, spec.ubu, spec.model_const, tol=1e-6, warm_ipm=1, condense=1)
    x0 = synth_states(n_samples, seed).numpy()
    o = cp.unit(1, pd, 0, 30, spec.p_nominal, x0, threads=threads)  # converge (untimed)
    it = o["iterate"]
    rng = np.random.default_rng(seed)
    ts = []
    x1 = x0
    for s in range(warmup + steps):
        if workload == "cartpole_tiny_pert":
            x1 = x0 + 1e-3 * rng.standard_normal(x0.shape)
        else:  # one environment step under the policy of the previous solve
            x1 = env_step_np(x1, o["u0"][:, 0])
        t0 = time.perf_counter()
        o = cp.unit(1, pd, 0, 1, spec.p_nominal, x1, iterate=it, threads=threads)
        dt = time.perf_counter() - t0
        it = o["iterate"]
        if s >= warmup:
            ts.append(dt)
    nthreads = cp.lib().cpu_port_get_threads() if threads == 0 else threads
    return (n_samples * len(ts) / sum(ts), nthreads, sum(ts) / len(ts),
            "restated structure-exploiting SQP-RTI + adjoint sensitivities in C++ (oracle/cpu_port), NOT acados")


def cpu_literal_run(n_samples, seed=1234):
    """What the reference really does per sample: a dense Python/torch restatement of
    update_nlp -- dense dR/dz (nz = 540) + SuperLU with ntheta right-hand sides -- after one dense SQP step."""
    import torch

    from oracle.problems import make_cartpole
    from oracle.solver import DenseSolver

    torch.set_num_threads(1)
    s = DenseSolver(make_cartpole("original"))
    x0 = synth_states(n_samples, seed).numpy()
    sols = [s.solve(x, tol=1e-8) for x in x0[:n_samples]]  # untimed: converged iterates
    t0 = time.perf_counter()
    for x, sol in zip(x0, sols):
        x1 = env_step_np(x[None, :], np.array([sol.U[0, 0]]))[0]
        s.unit(x1, init=(sol.U, sol.X), max_iter=1, polish=False)
    dt = time.perf_counter() - t0
    return n_samples / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    n = 16384 if not args.workload.startswith("chain_mass") else 32
    val, cores, sec, what = cpu_port_run(args.workload, n, args.steps, args.warmup)
    B = args.batch or wl["batch"]
    line = {"impl": "reference", "metric": wl["metric"], "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(B, args.gpus, args.workload, args.scaling),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} samples of the same workload per step (a bounded sample of the {B}-sample batch, "
                                       f"rate-based); {what}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch

    from mpc4rl_b200 import BatchedMPC

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    B = args.batch or (wl["batch"] // world if args.scaling == "strong" else wl["batch"])
    n_steps = args.warmup + args.steps
    W = Workload(args.workload, B, rank, dev)
    spec = W.spec
    mpc = BatchedMPC(spec, max_batch=B, device=local_rank)
    chain = args.workload.startswith("chain_mass")
    mpc.set_option("tol", 1e-6)
    if args.graph:
        mpc.set_option("graph", 1)
    split_opt = 2.0  # library default; "--opt split=n" overrides
    for kv in args.opt:
        k, v = kv.split("=")
        mpc.set_option(k, float(v))
        if k == "split":
            split_opt = float(v)
    conv_frac = W.setup(mpc, n_steps)
    g = torch.Generator(device="cpu").manual_seed(199 + rank)
    td = torch.randn(B, generator=g, dtype=torch.float64).to(dev)
    out = mpc.alloc_outputs(B)
    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        mpc.solve_sens(W.advance(i, out, mpc), max_sqp=1, out=out)
        acc = mpc.td_grad(td, out["dL"], out["status"])
        if dist is not None:
            dist.all_reduce(acc)
    barrier()
    l0 = mpc.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(4)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clk:
        ev[0].record(stream)
        for i in range(args.steps):
            x = W.advance(args.warmup + i, out, mpc)
            kev[i][0].record(stream)
            mpc.solve_sens(x, max_sqp=1, out=out)
            kev[i][1].record(stream)
            acc = mpc.td_grad(td, out["dL"], out["status"])
            kev[i][2].record(stream)
            if dist is not None:
                dist.all_reduce(acc)
            kev[i][3].record(stream)
            ev[i + 1].record(stream)
        barrier()
        total_ms = ev[0].elapsed_time(ev[-1])
        time.sleep(0.25)
    launches = mpc.launch_count - l0
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b, _, _ in kev]))
    allreduce_ms = float(np.mean([c.elapsed_time(d) for _, _, c, d in kev]))
    ok_frac = float((out["status"] == 0).double().mean().item())
    rmax = out["res"].max(dim=1).values
    res_max, res_med = float(rmax.max().item()), float(rmax.median().item())
    res_small = float((rmax < 1e-3).double().mean().item())

    # ---- per-kernel device times: replay of the first timed steps with the library's phase events, un-split ----
    phase_ms, queue, n_ph = {}, {}, min(3, args.steps)
    mpc.set_option("timing", 1)
    mpc.set_option("split", 1)
    W.rewind(mpc)
    for i in range(args.warmup + n_ph):
        x = W.logged(i) if W.store is None else W.advance(i, out, mpc, record=False)
        mpc.solve_sens(x, max_sqp=1, out=out)
        if i >= args.warmup:
            for k, v in mpc.timings().items():
                if k.startswith("queue"):
                    queue[k] = queue.get(k, 0.0) + v / n_ph
                else:
                    phase_ms[k] = phase_ms.get(k, 0.0) + v / n_ph
    if not chain:
        mpc.set_option("split", split_opt)
    mpc.set_option("timing", 0)

    # ---- e2e: host buffers through the C ABI, every step; the same states, replayed ----
    xs_host = [W.logged(i).cpu().pin_memory().numpy() for i in range(n_steps)]
    host_out = mpc.alloc_host_outputs(B, pinned=True)
    W.rewind(mpc)
    for i in range(args.warmup):
        if W.store is not None:
            W.store.load(W.idx[i])
        mpc.solve_sens_host(xs_host[i], max_sqp=1, out=host_out)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        if W.store is not None:
            W.store.load(W.idx[args.warmup + i])
        mpc.solve_sens_host(xs_host[args.warmup + i], max_sqp=1, out=host_out)
    barrier()
    e2e_s = time.perf_counter() - t0
    ng, nu = mpc.ngrad, spec.nu
    h2d = B * spec.nx * 8
    d2h = B * ((nu + 1 + 4 + ng + nu * ng) * 8 + 4)

    # ---- K-to-convergence: cold solve (MPC.reset guess) + sensitivities of the whole batch ----
    W.install_guess(mpc)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    oc = mpc.solve_sens(W.x0, max_sqp=60, out=out)
    c1.record(stream)
    torch.cuda.synchronize()
    conv_ms = c0.elapsed_time(c1)
    conv_ok = float((oc["status"] == 0).double().mean().item())

    tt = torch.tensor([total_ms, e2e_s * 1e3, kernel_ms, allreduce_ms, conv_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kernel_ms, allreduce_ms, conv_ms = (float(v) for v in tt.tolist())
    if rank == 0:
        peaks, peak_src = measured_peaks()
        units = B * world * args.steps
        value = units / (total_ms * 1e-3)
        achieved = B * wl["b_alg"] / (kernel_ms * 1e-3) / 1e9
        fp64_peak = mpc.fp64_peak_tflops()
        fpu = FLOPS_PER_UNIT.get(args.workload)
        if fpu is not None:
            fpu += dmma_flops_per_unit(args.workload, queue.get("queue_ipm_iters", 0.0) / B)
        fp64 = {"peak": fp64_peak, "unit": "TFLOP/s", "peak_source": "measured in this run: rlmpc_fp64_peak (FMA probe kernel, FMA = 2 flop)"}
        if fpu is not None:
            fa = B * fpu / (kernel_ms * 1e-3) / 1e12
            fp64.update({"achieved": fa, "frac": fa / fp64_peak, "flops_per_unit": fpu, "flops_source": FLOPS_SOURCE})
        dom = max(phase_ms, key=phase_ms.get) if phase_ms else None
        line = {
            "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(B, world, args.workload, args.scaling, args.graph),
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"],
                         "traffic": TRAFFIC_PER_STEP.get(args.workload) if B == wl["batch"] else None,
                         "traffic_source": "profiles/r02z_step_*.csv (dram read + write bytes summed over the kernels of one step, ncu)",
                         "peak_source": peak_src,
                         "kernel": ("rlmpc_solve_sens = k_chain_stage<lin> + k_chain_qp + k_chain_stage<hess> + k_chain_sens + k_chain_param"
                                    if chain else "rlmpc_solve_sens = k_lin + k_qp1 + k_qp3 + k_sens_stage + k_sens_sweep") + f" (dominant: {dom})",
                         "kernel_ms": kernel_ms, "kernels_ms": {k: round(v, 4) for k, v in phase_ms.items()},
                         "algorithmic_bytes_per_unit": wl["b_alg"],
                         "fp64": fp64,
                         "note": "the path is FP64 issue / latency bound, not HBM bound (SURVEY.md 8(d)): the HBM fraction is reported "
                                 "as defined (algorithmic bytes, not padded), the FP64 fraction is the one that says how far there is to go"},
            "allreduce_ms": allreduce_ms,
            "converge": {"what": "cold start (MPC.reset guess), SQP to tol 1e-6 (max 60 iterations) + sensitivities, whole batch, one call",
                         "ms": conv_ms, "value": B * world / (conv_ms * 1e-3), "unit": UNIT, "status0_frac": conv_ok},
            # KKT residual of the iterate after the RTI step (one SQP iteration, so not converged by construction)
            "quality": {"status0_frac_after_setup": conv_frac, "status0_frac_last_step": ok_frac,
                        "kkt_res_median_last_step": res_med, "kkt_res_lt_1e-3_frac_last_step": res_small,
                        "kkt_res_max_last_step": res_max,
                        "queue_frac": queue.get("queue_len", 0.0) / B,
                        "queue_ipm_iters_mean": (queue.get("queue_ipm_iters", 0.0) / B if chain else
                                                 queue.get("queue_ipm_iters", 0.0) / max(queue.get("queue_len", 0.0), 1.0))},
        }
        if world == 1 and not args.no_cpu:
            n_cpu = args.cpu_samples if not chain else 32
            cval, cores, csec, what = cpu_port_run(args.workload if (chain or args.workload == "cartpole_tiny_pert") else "cartpole", n_cpu, 3, 1)
            line["cpu_baseline"] = {"value": cval, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_cpu} samples of the same workload x 3 steps ({csec:.2f} s per step); {what}"}
            if args.workload == "cartpole" and args.literal_samples > 0:
                lval, lsec = cpu_literal_run(args.literal_samples)
                line["cpu_baseline_literal"] = {
                    "value": lval, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"{args.literal_samples} samples ({lsec:.1f} s): one dense SQP step + the literal restatement of update_nlp "
                              "(oracle/nlp.py: dense 540 x 540 dR/dz via torch.func, SuperLU with 83 right-hand sides), single process -- "
                              "what the reference's Python loop does per sample (BASELINE.md section 4, B1), NOT acados"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (default: the workload's)")
    ap.add_argument("--workload", default="cartpole", choices=sorted(WORKLOADS), help="cartpole = BASELINE.json configs[1] (headline)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's batch per GPU; strong: the workload's batch split over the GPUs (SURVEY.md 8(e))")
    ap.add_argument("--cpu-samples", type=int, default=16384)
    ap.add_argument("--literal-samples", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the RTI kernel chain as a CUDA graph (engine option graph=1)")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (tuning experiments)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
