#!/usr/bin/env python
"""Small invocation of every kernel, for compute-sanitizer memcheck (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpt_b200  # noqa: E402
from dpt_b200 import kernels  # noqa: E402
from dpt_b200.models.net import Transformer  # noqa: E402

means, _, _ = kernels.bandit_sample_means(77, 5, 1, 0)
kernels.bandit_rollin(means, 52, 0.3, 1, 0, dump=True, stats=torch.zeros(3, dtype=torch.float64, device="cuda"))
kernels.bandit_rollin(means, 51, 0.3, 1, 0)                                    # generic path
m7, _, _ = kernels.bandit_sample_means(9, 7, 1, 0)
kernels.bandit_rollin(m7, 33, 0.3, 1, 0)
goals = torch.randint(0, 10, (45, 2), dtype=torch.int32, device="cuda")
kernels.darkroom_rollin(goals, 10, 100, "uniform", 3, 0, torch.randint(0, 120, (45,), dtype=torch.int32, device="cuda"), 2, dump=True)
kernels.darkroom_rollin(goals, 10, 23, "expert", 3, 0, None, 1)
arms = np.random.RandomState(1234).normal(size=(10, 2)) / np.sqrt(2)
m10, _, _ = kernels.bandit_sample_means(50, 10, 0, 0)
for kind, mm, par in (("opt", means, {}), ("emp", means, dict(p0=1.0)), ("ucb", means, dict(p0=1.0)),
                      ("thompson", means, dict(p0=0.3, p1=0.5, p2=1 / 12.0)), ("linucb", m10, dict(p0=1.0, arms=arms))):
    kernels.online_loop(kind, mm, 45, 0.3, 2, 0, dump=True, **par)
    kernels.online_loop(kind, mm, 33, 0.3, 2, 0, materialise=False, **par)
o = kernels.online_loop("emp", means, 40, 0.3, 2, 0, p0=1.0)
kernels.arm_stats(o["context_actions"], o["context_rewards"], 30)
torch.manual_seed(0)
t = Transformer({"horizon": 40, "state_dim": 1, "action_dim": 5, "n_layer": 2, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
for prec in (0, 1):
    t.precision = prec
    out = t.online_loop(means[:20], 40, 0.3, True, 1, 0, dump=True)
    x = {"query_states": torch.ones(20, 1, device="cuda"), "context_states": out["context_states"], "context_actions": out["context_actions"],
         "context_next_states": out["context_next_states"], "context_rewards": out["context_rewards"]}
    t(x)
t2 = Transformer({"horizon": 140, "state_dim": 2, "action_dim": 5, "n_layer": 2, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": False})
for prec in (0, 1):
    t2.precision = prec
    for T in (17, 140):
        x = {"query_states": torch.rand(6, 2, device="cuda"), "context_states": torch.rand(6, T, 2, device="cuda"), "context_actions": torch.rand(6, T, 5, device="cuda"),
             "context_next_states": torch.rand(6, T, 2, device="cuda"), "context_rewards": torch.rand(6, T, 1, device="cuda")}
        t2(x)
kernels.darkroom_policy_rollout(torch.rand(45, 100, 5, device="cuda"), goals, 10, 30, True, 1, 0, 0, None, None, True)
torch.cuda.synchronize()
print("sanitize smoke done")
