"""One classical online-loop configuration, a few passes (target for ncu launch lists).  GPU box only.
    python scripts/ol_one.py emp 100000 500 5 [passes]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dpt_b200
from dpt_b200 import kernels

PAR = {"opt": {}, "emp": {"p0": 1.0}, "ucb": {"p0": 1.0}, "thompson": {"p0": 0.3, "p1": 0.5, "p2": 1 / 12.0},
       "linucb": {"p0": 1.0}}

if __name__ == "__main__":
    kind, N, H, d = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    passes = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    par = dict(PAR[kind])
    if kind == "linucb":
        par["arms"] = torch.tensor(np.random.RandomState(1234).normal(size=(d, 2)) / np.sqrt(2), dtype=torch.float64, device="cuda")
    means, _, _ = kernels.bandit_sample_means(N, d, 0, 0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(passes):
        a.record()
        out = kernels.online_loop(kind, means, H, 0.3, 2, 0, **par)
        b.record()
        torch.cuda.synchronize()
        print("pass %d: %.4f ms" % (i, a.elapsed_time(b)), flush=True)
