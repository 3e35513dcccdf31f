"""Multi-GPU layer: one process per GPU (torchrun), contiguous env-range sharding, no data-path
collective.

Envs are independent, so rank r of R owns the global env ids
``[r*N/R, (r+1)*N/R)``; Philox counters use the GLOBAL env id, hence every rank's output is
bit-identical to the corresponding slice of a single-GPU run.  The only exchange is the small
statistics block -- ``[3]`` return statistics for collection, ``[H,4]`` regret sums for the online
loop -- all-gathered / all-reduced over NCCL (NVLink 5 / NVSwitch on the B200 box; gloo in the CPU
tests).  Context tensors and K/V caches never leave the GPU that produced them.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total, rank, world_size):
    """Contiguous, balanced split of ``range(n_total)``: returns (lo, hi) of ``rank``."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_stats(local, group=None):
    """All-gather a small per-rank statistics tensor -> [world, ...] (same on every rank)."""
    rank, ws = world()
    if ws == 1:
        return local[None]
    out = torch.empty((ws,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out


def all_reduce_sums(local, group=None):
    """Sum a per-rank block of additive statistics over all ranks (in place, returns it)."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
    return local


def merge_return_stats(per_rank, steps_per_rank):
    """[world,3] (sum r, sum r^2, #optimal pulls) + env-steps per rank -> dict of global statistics."""
    per_rank = np.asarray(per_rank, dtype=np.float64)
    n = float(np.sum(steps_per_rank))
    s = per_rank.sum(0)
    mean = s[0] / n
    return {"env_steps": n, "mean_reward": mean, "var_reward": max(s[1] / n - mean * mean, 0.0),
            "frac_optimal_arm": s[2] / n}


def regret_stats_from_sums(sums, n_envs):
    """[H,4] regret sums over ``n_envs`` envs -> the four curves of evals/eval_bandit.py:169-178:
    per-step regret mean / sem and cumulative regret mean / sem (scipy.stats.sem, ddof = 1)."""
    s = np.asarray(sums, dtype=np.float64)
    n = float(n_envs)

    def mean_sem(s1, s2):
        mean = s1 / n
        var = np.maximum(s2 - s1 * s1 / n, 0.0) / max(n - 1.0, 1.0)
        return mean, np.sqrt(var / n)
    m, se = mean_sem(s[:, 0], s[:, 1])
    cm, cse = mean_sem(s[:, 2], s[:, 3])
    return {"mean": m, "sem": se, "regret_mean": cm, "regret_sem": cse}


# ------------------------------------------------------------------ sharded entry points -------
def collect_bandit_sharded(n_envs_total, dim, horizon, var, seed, device=None):
    """BASELINE config 5: bandit collection over all ranks.  Returns (local batch dict, global
    statistics dict).  The local batch holds this rank's env slice only."""
    from . import kernels
    rank, ws = world()
    lo, hi = shard_range(n_envs_total, rank, ws)
    means, opt_idx, opt_a = kernels.bandit_sample_means(hi - lo, dim, seed, lo, device)
    stats = torch.zeros(3, dtype=torch.float64, device=means.device)
    batch = kernels.bandit_rollin(means, horizon, float(var), seed, lo, stats=stats)
    batch.update(means=means, opt_a_index=opt_idx, optimal_actions=opt_a, env_range=(lo, hi))
    gathered = all_gather_stats(stats).cpu().numpy()
    steps = [(shard_range(n_envs_total, r, ws)[1] - shard_range(n_envs_total, r, ws)[0]) * horizon for r in range(ws)]
    return batch, merge_return_stats(gathered, steps)


def online_eval_sharded(kind, n_envs_total, dim, horizon, var, seed, model=None, materialise=False, **ctrl):
    """Online in-context evaluation over all ranks: each rank draws and evaluates its env slice with
    the fused loop, then the [H,4] regret sums are all-reduced.  ``kind``: 'opt' | 'emp' | 'ucb' |
    'thompson' | 'linucb' | 'transformer'.  Returns (local result dict, global regret curves)."""
    from . import kernels
    rank, ws = world()
    lo, hi = shard_range(n_envs_total, rank, ws)
    means, _, _ = kernels.bandit_sample_means(hi - lo, dim, seed, lo)
    if kind == "transformer":
        out = model.online_loop(means, horizon, var, ctrl.get("sample", True), seed, lo, materialise, True)
    else:
        out = kernels.online_loop(kind, means, horizon, var, seed, lo, materialise=materialise, regret=True, **ctrl)
    sums = all_reduce_sums(out["regret_sums"].clone())
    return out, regret_stats_from_sums(sums.cpu().numpy(), n_envs_total)
