#!/usr/bin/env python
"""Run `n` timed-region steps of a bench.py workload between cudaProfilerStart/Stop, so that
`ncu --profile-from-start off ...` sees exactly the kernels of the steps (not the setup solve).

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... \
        --clock-control none --csv --log-file profiles/<name>.csv python tools/profile_step.py --workload cartpole --steps 2
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mpc4rl_b200 import BatchedMPC  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cartpole")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
dev = torch.device("cuda", 0)
B = a.batch or bench.WORKLOADS[a.workload]["batch"]
W = bench.Workload(a.workload, B, 0, dev)
mpc = BatchedMPC(W.spec, max_batch=B, device=0)
mpc.set_option("tol", 1e-6)
for kv in a.opt:
    k, v = kv.split("=")
    mpc.set_option(k, float(v))
W.setup(mpc, a.warmup + a.steps)
out = mpc.alloc_outputs(B)
td = torch.randn(B, dtype=torch.float64, device=dev)
for i in range(a.warmup):
    mpc.solve_sens(W.advance(i, out, mpc), max_sqp=1, out=out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for i in range(a.steps):
    mpc.solve_sens(W.advance(a.warmup + i, out, mpc), max_sqp=1, out=out)
    mpc.td_grad(td, out["dL"], out["status"])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", a.steps, "steps of", a.workload, "batch", B, "status0", float((out["status"] == 0).double().mean()))
