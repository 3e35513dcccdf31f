"""CPU tier: the multi-GPU host logic with world_size 2 over gloo (no GPU): sharding, the statistics
exchange, and shard-independence of the Philox-addressed draws (restated by the oracle)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    import dpt_b200
    from dpt_b200.dist import shard_range
    for n in (0, 1, 7, 1000, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_regret_stats_from_sums_matches_oracle():
    import dpt_b200
    from dpt_b200.dist import regret_stats_from_sums
    from oracle import dpt_oracle as O
    rs = np.random.RandomState(0)
    opt, alg = rs.rand(40, 25) + 0.5, rs.rand(40, 25)
    diff = opt - alg
    cr = np.cumsum(diff, axis=1)
    sums = np.stack([diff.sum(0), (diff ** 2).sum(0), cr.sum(0), (cr ** 2).sum(0)], 1)
    got = regret_stats_from_sums(sums, 40)
    m, s, cm, cs = O.regret_stats(opt, alg)
    for a, b in ((got["mean"], m), (got["sem"], s), (got["regret_mean"], cm), (got["regret_sem"], cs)):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-12)


def _worker(rank, world_size, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    import dpt_b200
    from dpt_b200 import dist as D
    from oracle import dpt_oracle as O
    from oracle import philox as P
    N, H, d, seed = 37, 12, 5, 9
    lo, hi = D.shard_range(N, rank, world_size)
    ids = np.arange(lo, hi)
    # this rank's slice of the task + noise, addressed by GLOBAL env id (what the kernels do on the GPU)
    means = P.bandit_means(seed, ids, d)
    k = P.rollin_step_k(seed, ids, H)
    cov_idx, rand_idx = P.rollin_setup_ints(seed, ids, d)
    probs = np.full((hi - lo, d), 1.0 / d)
    z = np.zeros((hi - lo, H))
    xs, us, xps, rs, acts = O.rollin_bandit_batch(means, 0.3, cov_idx, probs, rand_idx, k * 2.0 ** -31, z)
    opt = means.argmax(1)
    local = torch.tensor([rs.sum(), (rs ** 2).sum(), float((acts == opt[:, None]).sum())], dtype=torch.float64)
    gathered = D.all_gather_stats(local)
    reg = (means.max(1)[:, None] - np.take_along_axis(means, acts, 1)).T            # [H, n_local]
    cr = np.cumsum(reg, axis=0)
    sums = torch.tensor(np.stack([reg.sum(1), (reg ** 2).sum(1), cr.sum(1), (cr ** 2).sum(1)], 1))
    sums = D.all_reduce_sums(sums)
    q.put((rank, lo, hi, acts, gathered.numpy(), sums.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    import dpt_b200
    from dpt_b200 import dist as D
    from oracle import dpt_oracle as O
    from oracle import philox as P
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N, H, d, seed = 37, 12, 5, 9
    ids = np.arange(N)
    means = P.bandit_means(seed, ids, d)
    cov_idx, rand_idx = P.rollin_setup_ints(seed, ids, d)
    xs, us, xps, rs, acts = O.rollin_bandit_batch(means, 0.3, cov_idx, np.full((N, d), 1.0 / d), rand_idx,
                                                  P.rollin_step_k(seed, ids, H) * 2.0 ** -31, np.zeros((N, H)))
    assert np.array_equal(np.concatenate([r[3] for r in res]), acts)                 # shards == slices of the whole
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == N
    for r in res:                                                                    # every rank holds the same gather
        assert np.array_equal(r[4], res[0][4]) and np.array_equal(r[5], res[0][5])
    st = D.merge_return_stats(res[0][4], [(r[2] - r[1]) * H for r in res])
    assert st["env_steps"] == N * H and abs(st["mean_reward"] - rs.mean()) < 1e-12
    assert abs(st["frac_optimal_arm"] - (acts == means.argmax(1)[:, None]).mean()) < 1e-12
    reg = means.max(1)[:, None] - np.take_along_axis(means, acts, 1)
    got = D.regret_stats_from_sums(res[0][5], N)
    m, s, cm, cs = O.regret_stats(np.repeat(means.max(1)[:, None], H, 1), np.take_along_axis(means, acts, 1))
    assert np.allclose(got["mean"], m) and np.allclose(got["sem"], s) and np.allclose(got["regret_mean"], cm) and np.allclose(got["regret_sem"], cs)


def test_collect_darkroom_sharded_needs_cuda():
    """No CPU fallback on the sharded darkroom collection either (it fails loudly without the CUDA path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import dpt_b200
    from dpt_b200 import dist as D
    with pytest.raises(dpt_b200._lib.DptError):
        D.collect_darkroom_sharded(np.zeros((8, 2), dtype=np.int64), 10, 4, 0)
