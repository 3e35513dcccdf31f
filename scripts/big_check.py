"""Full-size (C5: 1M envs x H=500 on ONE GPU, 16 GB) indexing check: the last shard of the big launch must equal a
separate 125k-env launch with the shifted global env id.  GPU box only."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dpt_b200
from dpt_b200 import kernels

N, H, d, seed = 1_000_000, 500, 5, 9
means, _, _ = kernels.bandit_sample_means(N, d, seed, 0)
big = kernels.bandit_rollin(means, H, 0.3, seed, 0)
lo = 875_000
part = kernels.bandit_rollin(means[lo:].contiguous(), H, 0.3, seed, lo)
for k in part:
    assert torch.equal(big[k][lo:], part[k]), k
assert float(big["context_actions"].sum()) == N * H
del big, part
torch.cuda.empty_cache()
# online loop at 4M envs x 128 steps (context 6.1 GB): last 1000 envs against a separate launch
N2, H2 = 4_000_000, 128
means2, _, _ = kernels.bandit_sample_means(N2, d, seed, 0)
o = kernels.online_loop("ucb", means2, H2, 0.3, seed, 0, p0=1.0)
p = kernels.online_loop("ucb", means2[-1000:].contiguous(), H2, 0.3, seed, N2 - 1000, p0=1.0)
assert torch.equal(o["context_actions"][-1000:], p["context_actions"]) and torch.equal(o["cum_means"][:, -1000:], p["cum_means"])
print("big_check ok")
