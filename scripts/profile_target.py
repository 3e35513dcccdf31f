#!/usr/bin/env python
"""Run ONE kernel target a few times (for ncu captures): python scripts/profile_target.py <target>"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpt_b200  # noqa: E402
from dpt_b200 import kernels  # noqa: E402

t = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if t == "darkroom":
    goals = torch.randint(0, 10, (100000, 2), dtype=torch.int32, device="cuda")
    for i in range(reps):
        kernels.darkroom_rollin(goals, 10, 100, "uniform", i, 0, None, 1)
elif t == "online1m_opt":
    means, _, _ = kernels.bandit_sample_means(1000000, 5, 0, 0)
    for i in range(reps):
        kernels.online_loop("opt", means, 100, 0.3, i, 0)
elif t.startswith("online_"):
    kind = t.split("_")[1]
    par = {"opt": {}, "emp": dict(p0=1.0), "ucb": dict(p0=1.0), "thompson": dict(p0=0.3, p1=0.5, p2=1 / 12.0),
           "linucb": dict(p0=1.0, arms=np.random.RandomState(1234).normal(size=(10, 2)) / np.sqrt(2))}[kind]
    d, H = (10, 200) if kind == "linucb" else (5, 500)
    means, _, _ = kernels.bandit_sample_means(100000, d, 0, 0)
    for i in range(reps):
        kernels.online_loop(kind, means, H, 0.3, i, 0, **par)
elif t in ("gpt2_c4", "gpt2_bf16_c4", "gpt2_c4_8gpu", "gpt2_bf16_c4_8gpu"):   # BASELINE configs[3]: 10k envs x H=500 (and its 8-GPU strong-scaled shard)
    from dpt_b200.models.net import Transformer
    torch.manual_seed(0)
    H = 500
    m = Transformer({"horizon": H, "state_dim": 1, "action_dim": 5, "n_layer": 4, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
    m.precision = 1 if "bf16" in t else 0
    means, _, _ = kernels.bandit_sample_means(int(os.environ.get("DPT_PROFILE_ENVS", 1250 if t.endswith("8gpu") else 10000)), 5, 0, 0)
    for i in range(reps):
        m.online_loop(means, H, 0.3, True, i, 0)
elif t in ("gpt2", "gpt2_bf16"):
    from dpt_b200.models.net import Transformer
    torch.manual_seed(0)
    H = 200
    m = Transformer({"horizon": H, "state_dim": 1, "action_dim": 5, "n_layer": 4, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
    m.precision = 1 if t == "gpt2_bf16" else 0
    means, _, _ = kernels.bandit_sample_means(4000, 5, 0, 0)
    for i in range(reps):
        m.online_loop(means, H, 0.3, True, i, 0)
elif t in ("dense_bf16", "dense_fp32"):
    from dpt_b200.models.net import Transformer
    torch.manual_seed(0)
    m = Transformer({"horizon": 100, "state_dim": 2, "action_dim": 5, "n_layer": 4, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
    m.precision = 1 if t == "dense_bf16" else 0
    B, T = 4096, 100
    x = {"query_states": torch.rand(B, 2, device="cuda"), "context_states": torch.rand(B, T, 2, device="cuda"), "context_actions": torch.rand(B, T, 5, device="cuda"),
         "context_next_states": torch.rand(B, T, 2, device="cuda"), "context_rewards": torch.rand(B, T, 1, device="cuda")}
    for i in range(reps):
        m(x)
elif t == "bandit":
    means, _, _ = kernels.bandit_sample_means(125000, 5, 0, 0)
    out = kernels.bandit_rollin(means, 500, 0.3, 0, 0)
    for i in range(reps):
        kernels.bandit_rollin(means, 500, 0.3, i, 0, out=out)
torch.cuda.synchronize()
print("done", t)
