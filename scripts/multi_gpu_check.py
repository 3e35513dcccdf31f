#!/usr/bin/env python
"""Multi-GPU check, run under torchrun on the GPU box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py
Every rank runs its env shard; rank 0 also runs the WHOLE problem on its own GPU and checks that each
shard is bit-identical to the matching slice (Philox counters use global env ids), and that the
NCCL-gathered statistics equal the single-GPU ones."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpt_b200  # noqa: E402
from dpt_b200 import dist as D, kernels  # noqa: E402


def digest(t):
    v = t.contiguous().view(torch.int32).long()
    w = torch.arange(1, v.numel() + 1, device=v.device, dtype=torch.long) % 1000003
    return torch.stack([v.sum(), (v.flatten() * w).sum()])


def main():
    rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    N, d, H, var, seed = 40000, 5, 100, 0.3, 123
    batch, stats = D.collect_bandit_sharded(N, d, H, var, seed)
    lo, hi = batch["env_range"]
    mine = torch.cat([digest(batch[k]) for k in ("context_actions", "context_rewards", "context_states")])
    allh = D.all_gather_stats(mine)
    # the same collection with the all-gather fused into the kernel (NVLink peer stores, no NCCL call)
    pg = D.PeerGather(slots=2)
    batch2, stats2 = D.collect_bandit_sharded(N, d, H, var, seed, peer=pg, peer_slot=1)
    p2p_ok = all(abs(stats2[k] - stats[k]) < 1e-12 for k in stats) and all(torch.equal(batch2[k], batch[k]) for k in ("context_actions", "context_rewards"))
    flag = torch.tensor([1.0 if p2p_ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    p2p_ok = bool(flag.item() == 1.0)
    # fewer envs than ranks: the ranks with an empty shard launch nothing and publish zeros (no hang, same totals)
    b3, stats3 = D.collect_bandit_sharded(1, d, H, var, seed, peer=pg, peer_slot=0)
    m1, _, _ = kernels.bandit_sample_means(1, d, seed, 0)
    st1 = torch.zeros(3, dtype=torch.float64, device="cuda")
    kernels.bandit_rollin(m1, H, var, seed, 0, stats=st1)
    tiny_ok = abs(stats3["mean_reward"] - float(st1[0]) / H) < 1e-12 and stats3["env_steps"] == H
    pg.close()
    # transformer controller (config 4 shape, small): sharded K/V-cached loop == slices of the single-GPU loop
    from dpt_b200.models.net import Transformer
    torch.manual_seed(0)
    net = Transformer({"horizon": 48, "state_dim": 1, "action_dim": d, "n_layer": 2, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
    NT = 3000
    tout, tcurves = D.online_eval_sharded("transformer", NT, d, 48, var, seed, model=net)
    allt = D.all_gather_stats(digest(tout["cum_means"]))
    out, curves = D.online_eval_sharded("thompson", N, d, H, var, seed, p0=var, p1=0.5, p2=1 / 12.0)
    cm = digest(out["cum_means"])
    allc = D.all_gather_stats(cm)
    # darkroom collection (config 2) sharded the same way
    goals = np.random.RandomState(5).randint(0, 10, (N, 2))
    dk, dstats = D.collect_darkroom_sharded(goals, 10, 64, seed)
    alld = D.all_gather_stats(torch.cat([digest(dk[k]) for k in ("context_states", "context_actions", "context_next_states", "context_rewards")]))
    ok = True
    if rank == 0:
        means, _, _ = kernels.bandit_sample_means(N, d, seed, 0)
        st = torch.zeros(3, dtype=torch.float64, device="cuda")
        full = kernels.bandit_rollin(means, H, var, seed, 0, stats=st)
        fo = kernels.online_loop("thompson", means, H, var, seed, 0, materialise=False, p0=var, p1=0.5, p2=1 / 12.0)
        for r in range(ws):
            a, b = D.shard_range(N, r, ws)
            want = torch.cat([digest(full[k][a:b]) for k in ("context_actions", "context_rewards", "context_states")])
            ok &= bool(torch.equal(want, allh[r]))
            ok &= bool(torch.equal(digest(fo["cum_means"][:, a:b]), allc[r]))
        fd = kernels.darkroom_rollin(goals, 10, 64, "uniform", seed, 0, None, 1)
        for r in range(ws):
            a, b = D.shard_range(N, r, ws)
            want = torch.cat([digest(fd[k][a:b]) for k in ("context_states", "context_actions", "context_next_states", "context_rewards")])
            ok &= bool(torch.equal(want, alld[r]))
        ok &= dstats["env_steps"] == N * 64 and abs(dstats["mean_reward"] - float(fd["context_rewards"].double().mean())) < 1e-12
        ref = D.merge_return_stats(st.cpu().numpy()[None], [N * H])
        ok &= abs(ref["mean_reward"] - stats["mean_reward"]) < 1e-9 and abs(ref["frac_optimal_arm"] - stats["frac_optimal_arm"]) < 1e-12
        rc = D.regret_stats_from_sums(fo["regret_sums"].cpu().numpy(), N)
        ok &= bool(np.allclose(rc["regret_mean"], curves["regret_mean"], rtol=1e-9)) and bool(np.allclose(rc["sem"], curves["sem"], rtol=1e-6))
        ok &= p2p_ok and tiny_ok
        tm, _, _ = kernels.bandit_sample_means(NT, d, seed, 0)
        tf = net.online_loop(tm, 48, var, True, seed, 0, False, True)
        for r in range(ws):
            a, b = D.shard_range(NT, r, ws)
            ok &= bool(torch.equal(digest(tf["cum_means"][:, a:b]), allt[r]))
        trc = D.regret_stats_from_sums(tf["regret_sums"].cpu().numpy(), NT)
        ok &= bool(np.allclose(trc["regret_mean"], tcurves["regret_mean"], rtol=1e-9))
        print("multi_gpu_check world=%d: %s [p2p gather %s] (mean reward %.5f, final cumulative regret %.3f +- %.3f)" % (
            ws, "OK" if ok else "MISMATCH", "OK" if p2p_ok else "MISMATCH", stats["mean_reward"], curves["regret_mean"][-1], curves["regret_sem"][-1]))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
