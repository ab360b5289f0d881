"""GPU debugging aid: the headline workload's step with the warp-per-sample queue kernel (coop=1) against the
thread-per-sample path (coop=0) from the SAME iterates; samples whose status differs are dumped for a host replay."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mpc4rl_b200 import BatchedMPC

dev = torch.device("cuda", 0)
B = 65536
W = bench.Workload("cartpole", B, 0, dev)
mpc = BatchedMPC(W.spec, max_batch=B, device=0)
mpc.set_option("tol", 1e-6)
W.setup(mpc, 8)
out = mpc.alloc_outputs(B)
for i in range(5):
    mpc.solve_sens(W.advance(i, out, mpc), max_sqp=1, out=out)
x = W.advance(5, out, mpc).clone()
store = mpc.iterate_store(B)
idx = torch.arange(B, dtype=torch.int32, device=dev)
store.save(idx)
res = {}
for coop in (1, 0):
    store.load(idx)
    mpc.set_option("coop", coop)
    o = mpc.solve_sens(x, max_sqp=1)
    res[coop] = {k: v.clone() for k, v in o.items()}
    print("coop", coop, "status counts", torch.bincount(o["status"].long(), minlength=5).tolist())
a, b = res[1], res[0]
diff = (a["status"] != b["status"])
print("status differs on", int(diff.sum()), "| coop bad & thread ok:", int(((a["status"] != 0) & (b["status"] == 0)).sum()),
      "| thread bad & coop ok:", int(((a["status"] == 0) & (b["status"] != 0)).sum()))
ok = (a["status"] == 0) & (b["status"] == 0)
print("max |du0| on common ok:", float((a["u0"] - b["u0"])[ok].abs().max()))
sel = torch.where((a["status"] != 0) & (b["status"] == 0))[0][:64]
if len(sel):
    it_size = store.buf.numel() // B
    raw = store.buf.view(-1)
    # store layout: tiles of 32 iterates, element-major inside a tile (engine layout)
    its = []
    for s_ in sel.tolist():
        t, l = divmod(s_, 32)
        its.append(raw[t * it_size * 32 + l: (t + 1) * it_size * 32: 32].cpu().numpy())
    np.savez("gpurun_out/coop_fail_samples.npz", x=x[sel].cpu().numpy(), it=np.array(its), idx=sel.cpu().numpy(),
             u0_thread=b["u0"][sel].cpu().numpy(), u0_coop=a["u0"][sel].cpu().numpy(), st_coop=a["status"][sel].cpu().numpy())
    print("dumped", len(sel))
