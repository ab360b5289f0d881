#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric   MPC solves+sensitivities/sec: one unit = for one sample, one SQP-RTI step of the cartpole
         NMPC (N=40, nx=4, nu=1, config/cartpole_original.yaml) from the stored warm-start iterate
         PLUS dL/dtheta and dpi/dtheta at the new iterate (= one `update`/`q_update` + one
         `update_nlp` of the reference).  SURVEY.md 8(d).
step     one rlmpc_solve_sens call over a batch of 65 536 synthetic samples per GPU (weak scaling): five
         kernels (linearise | convergence test + fast QP | full interior point on the queued samples |
         exact-Hessian stage evaluation | factorisation + adjoint sweeps), followed by the TD-gradient
         accumulator kernel and, for N>1, its NCCL all-reduce.
value    whole-job units/s with inputs resident in HBM, CUDA-event timed, max over ranks.
e2e      same metric through the C ABI host entry point (rlmpc_solve_sens_host): pinned host buffers,
         H2D of the states and D2H of every result inside the timed region (wall clock, max over ranks).

`--impl reference` times the CPU path instead: acados/CasADi are not installable here, so this is the
oracle's host port of the same structure-exploiting algorithm (oracle/cpu_port, all host threads) --
"restated, not acados" (BASELINE.md section 3).
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
BATCH = 65536  # headline workload; algorithmic bytes per unit: 8*(2*524 + 4+1 + 1+1+3+3) + 4 = 8492


def synth_states(B, seed):
    """s~U(-1,1), s_dot~U(-2,2), theta~U(-pi,pi), theta_dot~U(-4,4)."""
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    lo = torch.tensor([-1.0, -2.0, -np.pi, -4.0], dtype=torch.float64)
    return lo + (-2.0 * lo) * torch.rand(B, 4, generator=g, dtype=torch.float64)


WORKLOADS = {
    # BASELINE.json configs[1] -- the headline
    "cartpole": dict(batch=65536, b_alg=8492, metric=METRIC,
                     text="cartpole_original N=40 nx=4 nu=1 ntheta=83 (3 with gradient), V-mode SQP-RTI (K=1) + "
                          "dL/dtheta + dpi/dtheta, warm-started from the converged iterate, states perturbed every step"),
    # SURVEY.md 8(d) config 2, secondary: config/cartpole.yaml (N=30, box bounds on all states) -- `--workload cartpole_bx`
    "cartpole_bx": dict(batch=65536, b_alg=14092, metric="MPC solves+sensitivities/sec (cartpole.yaml N=30 with state bounds, batch 65k)",
                        text="cartpole.yaml N=30 nx=4 nu=1, input and state bounds, V-mode SQP-RTI (K=1) + dL/dtheta + dpi/dtheta, "
                             "warm-started from the converged iterate, states perturbed every step"),
    # BASELINE.json configs[3] (secondary line, `--workload evaporation`): N=100, nu=3 (third input = slack), affine h rows
    "evaporation": dict(batch=32768, b_alg=45228, metric="MPC solves+sensitivities/sec (evaporation N=100, batch 32k)",
                        text="evaporation_process N=100 nx=2 nu=3 ntheta=60 (tracking-cost parameters), V-mode SQP-RTI "
                             "(K=1) + dL/dtheta + dpi/dtheta, warm-started from the converged iterate, states perturbed every step"),
}


def config_dict(B, n_gpus, workload="cartpole"):
    return {"workload": WORKLOADS[workload]["text"],
            "batch_per_gpu": B, "global_batch": B * n_gpus, "parallelism": f"dp{n_gpus} (batch shards, replicated theta)",
            "l2": "working set (iterate + stage scratch, > 1.5 GB per GPU) is larger than L2, no flush needed",
            "seed": 1234}


def make_workload(name, B, rank, dev):
    """(spec, x0 [B,nx] on dev, perturbation scale, function that installs the initial guess)"""
    import torch

    if name == "cartpole":
        from mpc4rl_b200 import cartpole_original_config, cartpole_spec

        spec = cartpole_spec(cartpole_original_config())
        x0 = synth_states(B, 1234 + rank).to(dev)
        return spec, x0, 1e-3, lambda mpc: mpc.reset(x0)
    if name == "cartpole_bx":
        from mpc4rl_b200 import cartpole_config, cartpole_spec

        spec = cartpole_spec(cartpole_config())
        x0 = synth_states(B, 1234 + rank).to(dev)
        x0[:, 1:] *= 0.5  # keep the start inside the state box (|s_dot| <= 10, |theta| <= 6.28, |theta_dot| <= 10)
        return spec, x0, 1e-3, lambda mpc: mpc.reset(x0)
    from mpc4rl_b200 import evaporation_spec

    spec = evaporation_spec(gamma=0.99)
    g = torch.Generator(device="cpu").manual_seed(7 + rank)  # X_2~U(25,40), P_2~U(49.7,70)
    lo, hi = torch.tensor([25.0, 49.7], dtype=torch.float64), torch.tensor([40.0, 70.0], dtype=torch.float64)
    x0 = (lo + (hi - lo) * torch.rand(B, 2, generator=g, dtype=torch.float64)).to(dev)

    def guess(mpc):  # every stage on the steady state (evaporation_process/acados.py:104-109)
        mpc.reset(B=B)
        xs = torch.tensor(spec.x_init, dtype=torch.float64, device=dev).repeat(B, 1)
        us = torch.tensor(spec.u_init, dtype=torch.float64, device=dev).repeat(B, 1)
        for k in range(spec.N + 1):
            mpc.put("x", k, xs)
        for k in range(spec.N):
            mpc.put("u", k, us)

    return spec, x0, 1e-2, guess


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
def cpu_port_run(n_samples, steps, warmup, threads=0, seed=1234):
    """RTI + sensitivities on the host (oracle/cpu_port): returns (units/s, threads, seconds per step)."""
    from oracle import cpu_port as cp
    from mpc4rl_b200.problems import cartpole_original_config, cartpole_spec

    spec = cartpole_spec(cartpole_original_config())
    pd = cp.make_pd(spec.N, spec.cost_scaling(), spec.lbu, spec.ubu, spec.model_const, tol=1e-6, warm_ipm=1, condense=1)
    x0 = synth_states(n_samples, seed).numpy()
    o = cp.unit(1, pd, 0, 30, spec.p_nominal, x0, threads=threads)  # converge (untimed)
    it = o["iterate"]
    rng = np.random.default_rng(seed)
    ts = []
    for s in range(warmup + steps):
        x1 = x0 + 1e-3 * rng.standard_normal(x0.shape)
        t0 = time.perf_counter()
        o = cp.unit(1, pd, 0, 1, spec.p_nominal, x1, iterate=it, threads=threads)
        dt = time.perf_counter() - t0
        it = o["iterate"]
        if s >= warmup:
            ts.append(dt)
    nthreads = cp.lib().cpu_port_get_threads() if threads == 0 else threads
    return n_samples * len(ts) / sum(ts), nthreads, sum(ts) / len(ts)


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = 16384
    val, cores, sec = cpu_port_run(n, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(BATCH, args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} samples of the same workload per step (bounded sample of the 65536 batch); "
                                       "restated structure-exploiting SQP-RTI + adjoint sensitivities in C++, NOT acados"},
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
    B = args.batch or wl["batch"]
    spec, x0, pert, install_guess = make_workload(args.workload, B, rank, dev)
    mpc = BatchedMPC(spec, max_batch=B, device=local_rank)
    # untimed setup: converge the batch once (cold start like MPC.reset), this is the warm-start state
    mpc.set_option("tol", 1e-6)
    mpc.set_option("timing", 1)
    split_opt = 2.0  # library default; "--opt split=n" overrides
    for kv in args.opt:
        k, v = kv.split("=")
        mpc.set_option(k, float(v))
        if k == "split":
            split_opt = float(v)
    install_guess(mpc)
    _, _, st = mpc.solve(x0, max_sqp=60)
    torch.cuda.synchronize()
    conv_frac = float((st == 0).double().mean().item())
    n_steps = args.warmup + args.steps
    g = torch.Generator(device="cpu").manual_seed(99 + rank)
    # every step sees fresh states: the converged ones moved by a small "environment step"
    xs = [(x0 + pert * torch.randn(B, spec.nx, generator=g, dtype=torch.float64).to(dev)) for _ in range(n_steps)]
    td = torch.randn(B, generator=g, dtype=torch.float64).to(dev)
    out = mpc.alloc_outputs(B)
    stream = torch.cuda.current_stream()

    def step(i):
        mpc.solve_sens(xs[i], max_sqp=1, out=out)
        acc = mpc.td_grad(td, out["dL"], out["status"])
        if dist is not None:
            dist.all_reduce(acc)
        return acc

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    l0 = mpc.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    phase_ms = {k: 0.0 for k in mpc.PHASES}
    with ClockSampler(local_rank) as clk:
        ev[0].record(stream)
        for i in range(args.steps):
            kev[i][0].record(stream)
            mpc.solve_sens(xs[args.warmup + i], max_sqp=1, out=out)
            kev[i][1].record(stream)
            acc = mpc.td_grad(td, out["dL"], out["status"])
            if dist is not None:
                dist.all_reduce(acc)
            ev[i + 1].record(stream)
        barrier()
        total_ms = ev[0].elapsed_time(ev[-1])
        time.sleep(0.25)
    launches = mpc.launch_count - l0
    # per-kernel device times (CUDA events recorded by the library on the launching stream between
    # the kernels of a call); separate pass because reading them waits for the call
    n_ph = 3
    queue = {}
    mpc.set_option("split", 1)  # per-kernel times of the un-split call (split parts overlap on two streams)
    for i in range(n_ph):
        mpc.solve_sens(xs[args.warmup + (i % args.steps)], max_sqp=1, out=out)
        for k, v in mpc.timings().items():
            if k in phase_ms:
                phase_ms[k] += v / n_ph
            else:
                queue[k] = queue.get(k, 0.0) + v / n_ph
    mpc.set_option("split", split_opt)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    ok_frac = float((out["status"] == 0).double().mean().item())
    rmax = out["res"].max(dim=1).values
    res_max = float(rmax.max().item())
    res_med = float(rmax.median().item())
    res_small = float((rmax < 1e-3).double().mean().item())

    # ---- e2e: host buffers through the C ABI, every step ----
    xs_host = [x.cpu().pin_memory().numpy() for x in xs]  # inputs in pinned host memory
    host_out = mpc.alloc_host_outputs(B, pinned=True)       # results read back into pinned host memory
    install_guess(mpc)
    mpc.solve(x0, max_sqp=60)
    for i in range(args.warmup):
        mpc.solve_sens_host(xs_host[i], max_sqp=1, out=host_out)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        o = mpc.solve_sens_host(xs_host[args.warmup + i], max_sqp=1, out=host_out)
    barrier()
    e2e_s = time.perf_counter() - t0
    ng, nu = mpc.ngrad, spec.nu
    h2d = B * spec.nx * 8
    d2h = B * ((nu + 1 + 4 + ng + nu * ng) * 8 + 4)

    tt = torch.tensor([total_ms, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kernel_ms = (float(v) for v in tt.tolist())
    if rank == 0:
        peaks, peak_src = measured_peaks()
        units = B * world * args.steps
        value = units / (total_ms * 1e-3)
        achieved = B * wl["b_alg"] / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(B, world, args.workload),
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of the kernels of one step, ncu launch list
                         # profiles/r01h_launches_rti_steps.csv (headline workload at its default batch only)
                         "traffic": 8.960e9 if (args.workload == "cartpole" and B == 65536) else None,
                         "traffic_source": "profiles/r01h_launches_rti_steps.csv (sum over the 7 kernels of one un-split call)",
                         "peak_source": peak_src,
                         "kernel": "rlmpc_solve_sens = k_lin + k_qp1 + k_qp3 + k_sens_stage + k_sens_sweep "
                                   f"(dominant: {max(phase_ms, key=phase_ms.get)})",
                         "kernel_ms": kernel_ms, "kernels_ms": {k: round(v, 4) for k, v in phase_ms.items()},
                         "algorithmic_bytes_per_unit": wl["b_alg"],
                         "note": "path is FP64 CUDA-core/latency bound, not HBM bound (SURVEY.md 8(d)); "
                                 "fraction reported as defined, bytes not padded; k_qp1 and k_sens_sweep run at "
                                 "70-100 % of the measured HBM peak on the bytes they actually move (profiles/r01_summary.md)"},
            # KKT residual of the iterate after the RTI step (one SQP iteration, so not converged by construction);
            # the max comes from the states on which full-step Gauss-Newton SQP 2-cycles (status 2 in the setup solve)
            "quality": {"status0_frac_after_setup": conv_frac, "status0_frac_last_step": ok_frac,
                        "kkt_res_median_last_step": res_med, "kkt_res_lt_1e-3_frac_last_step": res_small,
                        "kkt_res_max_last_step": res_max,
                        # samples handed to the full interior-point pass per step, and its mean iteration count
                        "queue_frac": queue.get("queue_len", 0.0) / B,
                        "queue_ipm_iters_mean": queue.get("queue_ipm_iters", 0.0) / max(queue.get("queue_len", 0.0), 1.0)},
        }
        if world == 1 and not args.no_cpu and args.workload == "cartpole":
            cval, cores, csec = cpu_port_run(args.cpu_samples, 3, 1)
            line["cpu_baseline"] = {"value": cval, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_samples} samples of the same workload x 3 steps ({csec:.2f} s per step); "
                                              "restated structure-exploiting C++ port (oracle/cpu_port), NOT acados"}
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
    ap.add_argument("--cpu-samples", type=int, default=16384)
    ap.add_argument("--no-cpu", action="store_true")
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
