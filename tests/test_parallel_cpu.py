"""Host-side data-parallel logic on CPU with gloo, world_size 2 (the kernels themselves need a GPU;
what is tested here is the sharding + accumulator all-reduce + identical parameter step on every rank)."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, nth, out_dir):
    import torch.distributed as dist

    from mpc4rl_b200.parallel import allreduce_accumulator, shard_range, td_parameter_step

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)  # every rank draws the same global batch, then takes its shard
    td = torch.randn(n, generator=g, dtype=torch.float64)
    dQ = torch.randn(n, nth, generator=g, dtype=torch.float64)
    status = (torch.rand(n, generator=g) < 0.1).to(torch.int32)  # 10 % failed solves are masked out
    lo, hi = shard_range(n, rank, world)
    ok = status[lo:hi] == 0
    acc = torch.zeros(nth + 2, dtype=torch.float64)  # what rlmpc_td_grad produces on each rank's shard
    acc[:nth] = (td[lo:hi][ok].unsqueeze(1) * dQ[lo:hi][ok]).sum(0)
    acc[nth] = td[lo:hi][ok].sum()
    acc[nth + 1] = ok.sum()
    allreduce_accumulator(acc)
    theta = torch.arange(nth + 3, dtype=torch.float64)
    new = td_parameter_step(theta, acc, lr=1e-2)
    torch.save({"acc": acc, "theta": new, "shard": (lo, hi)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_td_accumulator_allreduce(tmp_path):
    n, nth, world = 1001, 12, 2
    mp.spawn(_worker, args=(world, _free_port(), n, nth, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{i}.pt")) for i in range(world)]
    assert r[0]["shard"] == (0, 501) and r[1]["shard"] == (501, 1001)
    assert torch.equal(r[0]["acc"], r[1]["acc"]) and torch.equal(r[0]["theta"], r[1]["theta"])
    g = torch.Generator().manual_seed(0)
    td = torch.randn(n, generator=g, dtype=torch.float64)
    dQ = torch.randn(n, nth, generator=g, dtype=torch.float64)
    ok = (torch.rand(n, generator=g) < 0.1).to(torch.int32) == 0
    ref = (td[ok].unsqueeze(1) * dQ[ok]).sum(0)
    assert torch.allclose(r[0]["acc"][:nth], ref, rtol=1e-12, atol=1e-12)
    assert r[0]["acc"][nth + 1].item() == ok.sum().item()
    expect = torch.arange(nth + 3, dtype=torch.float64)
    expect[:nth] += 1e-2 * ref / ok.sum()
    assert torch.allclose(r[0]["theta"], expect, rtol=1e-12, atol=1e-12)


def test_shard_range_covers_everything():
    from mpc4rl_b200.parallel import shard_range

    for n in (0, 1, 7, 65536, 65537):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
