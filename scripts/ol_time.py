"""Classical online loop: GPU time per pass (CUDA events, outputs preallocated by the wrapper each pass) and host enqueue time.
    python scripts/ol_time.py kind N H d [passes]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dpt_b200
from dpt_b200 import kernels
from ol_one import PAR

if __name__ == "__main__":
    kind, N, H, d = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    passes = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    par = dict(PAR[kind])
    if kind == "linucb":
        par["arms"] = torch.tensor(np.random.RandomState(1234).normal(size=(d, 2)) / np.sqrt(2), dtype=torch.float64, device="cuda")
    means, _, _ = kernels.bandit_sample_means(N, d, 0, 0)
    for _ in range(3):
        kernels.online_loop(kind, means, H, 0.3, 2, 0, **par)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(passes + 1)]
    host = []
    ev[0].record()
    for i in range(passes):
        t0 = time.perf_counter()
        kernels.online_loop(kind, means, H, 0.3, 2, 0, **par)
        host.append(time.perf_counter() - t0)
        ev[i + 1].record()
    torch.cuda.synchronize()
    gpu = [ev[i].elapsed_time(ev[i + 1]) for i in range(passes)]
    # one isolated pass: sync before and after
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    kernels.online_loop(kind, means, H, 0.3, 2, 0, **par)
    b.record()
    torch.cuda.synchronize()
    print("%s N=%d H=%d d=%d: gpu back-to-back %.4f ms (min %.4f), isolated %.4f ms, host enqueue %.4f ms" % (
        kind, N, H, d, float(np.mean(gpu)), min(gpu), a.elapsed_time(b), 1e3 * float(np.mean(host))), flush=True)
