"""Online-loop kernel time against the env count (wave quantisation / latency floor study).  GPU box only."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dpt_b200
from dpt_b200 import kernels

PAR = {"opt": {}, "emp": {}, "ucb": {"p0": 1.0}, "thompson": {"p0": 0.3, "p1": 0.5, "p2": 1 / 12.0}}


REGRET = os.environ.get("OL_REGRET", "1") == "1"


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


if __name__ == "__main__":
    H, d = 500, 5
    for kind in sys.argv[1:] or ["opt", "emp", "ucb", "thompson"]:
        for N in [int(x) for x in os.environ.get("OL_SWEEP_N", "4736,18944,47360,75776,94720,100000,113664,189440,400000").split(",")]:
            means, _, _ = kernels.bandit_sample_means(N, d, 0, 0)
            for mat in (True, False):
                ms = timeit(lambda: kernels.online_loop(kind, means, H, 0.3, 2, 0, materialise=mat, regret=REGRET, **PAR[kind]), 8)
                print(json.dumps({"kind": kind, "N": N, "materialise": mat, "ms": round(ms, 4), "gsteps": round(N * H / ms / 1e6, 2)}), flush=True)
