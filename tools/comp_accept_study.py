import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from mpc4rl_b200 import BatchedMPC, cartpole_original_config, cartpole_spec
g = np.load("tests/golden/cartpole_original_rti.npz")
T = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda:0")
spec = cartpole_spec(cartpole_original_config())
B = g["x0"].shape[0]
for ca in (0.5, 0.2, 0.1, 0.05, 0.02):
    res = {}
    for path in ("warm", "cold"):
        mpc = BatchedMPC(spec, max_batch=B, device=0)
        mpc.set_option("comp_accept", ca)
        x0 = T(g["x0"])
        mpc.reset(x0)
        if path == "warm":
            mpc.set_option("tol", 1e-9)
            u0, _, st = mpc.solve(x0, max_sqp=300)
            sel = (g["status"] == 0) & (st.cpu().numpy() == 0) & (np.abs(u0.cpu().numpy() - g["u0"]).max(1) < 1e-6)
            mpc.set_option("tol", 1e-6)
        else:
            for k in range(spec.N + 1):
                mpc.put("x", k, T(g["X"][:, k]))
            for k in range(spec.N):
                mpc.put("u", k, T(g["U"][:, k]))
            sel = np.ones(B, bool)
        out = mpc.solve_sens(T(g["x1"]), max_sqp=1)
        sel = sel & (out["status"].cpu().numpy() == 0)
        du = np.abs(out["u0"].cpu().numpy() - g["u1"])[sel].max()
        dV = np.abs(out["cost"].cpu().numpy() - g["V1"])[sel].max() / np.abs(g["V1"][sel]).max()
        dL = np.abs(mpc.full_grad(out["dL"]).cpu().numpy() - g["dV1"])[sel].max() / np.abs(g["dV1"][sel]).max()
        dpi = np.abs(mpc.full_grad(out["dpi"]).cpu().numpy() - g["dpi1"])[sel].max() / np.abs(g["dpi1"][sel]).max()
        res[path] = (sel.mean(), du, dV, dL, dpi)
    print("comp_accept", ca, {k: tuple(float("%.3g" % x) for x in v) for k, v in res.items()}, flush=True)
