"""Cold-start throughput (what a replay-buffer sample without a stored iterate costs): MPC.reset-like
initial guess, SQP to convergence (max 30 iterations), V- and Q-mode, with the warp-per-sample and with the thread-per-sample queue kernel."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from bench import synth_states
from mpc4rl_b200 import BatchedMPC, cartpole_original_config, cartpole_spec

B = 65536
spec = cartpole_spec(cartpole_original_config())
x0 = synth_states(B, 1234).cuda()
u0 = (160.0 * torch.rand(B, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(1)) - 80.0).cuda()
for cond in (1, 0):
    m = BatchedMPC(spec, max_batch=B, device=0)
    m.set_option("coop", cond)  # 1: warp-per-sample queue kernel, 0: thread-per-sample (partially condensed)
    for mode, a in (("V", None), ("Q", u0)):
        ts = []
        for rep in range(3):
            m.reset(x0)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = m.solve_sens(x0, a, max_sqp=30)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        ok = (out["status"] == 0).double().mean().item()
        print(f"coop={cond} mode={mode}: {min(ts)*1e3:.1f} ms per 65536 cold solves+sens ({B/min(ts)/1e6:.2f} M units/s), converged {ok:.3f}")
