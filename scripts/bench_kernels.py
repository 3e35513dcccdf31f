#!/usr/bin/env python
"""Per-kernel throughput table (CUDA events, warm-up, outputs > L2 where the config allows):
env-steps/s, algorithmic GB/s and fraction of the measured HBM peak for every kernel of the path.
Writes one JSON object per line to stdout; run on the GPU box.

    python scripts/bench_kernels.py [--reps 20]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpt_b200  # noqa: E402
from dpt_b200 import kernels  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    return float(np.mean(per)), float(np.min(per))


def report(name, steps, bytes_per_step, ms_mean, ms_min, **extra):
    gbs = steps * bytes_per_step / (ms_mean * 1e-3) / 1e9
    print(json.dumps(dict(kernel=name, env_steps=steps, ms_mean=ms_mean, ms_min=ms_min, env_steps_per_s=steps / (ms_mean * 1e-3),
                          algorithmic_bytes_per_step=bytes_per_step, achieved_gbs=gbs, frac_of_measured_hbm_peak=gbs / PEAK,
                          **extra)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    R = a.reps
    want = lambda n: not a.only or a.only in n   # noqa: E731

    if want("bandit_rollin"):
        for N, H, d in [(125000, 500, 5), (1000, 500, 5), (100000, 200, 10)]:
            means, _, _ = kernels.bandit_sample_means(N, d, 0, 0)
            out = kernels.bandit_rollin(means, H, 0.3, 0, 0)
            m, mn = timeit(lambda: kernels.bandit_rollin(means, H, 0.3, 1, 0, out=out), R)
            report("bandit_rollin N=%d H=%d d=%d" % (N, H, d), N * H, 4 * (3 + d), m, mn)
            if N == 125000:   # bench.py's step: the same launch also accumulating the return statistics
                st = torch.zeros(3, dtype=torch.float64, device="cuda")
                m, mn = timeit(lambda: kernels.bandit_rollin(means, H, 0.3, 1, 0, out=out, stats=st), R)
                report("bandit_rollin N=%d H=%d d=%d + return statistics" % (N, H, d), N * H, 4 * (3 + d), m, mn)
            del out
    if want("darkroom"):
        for N, H in [(100000, 100), (1000000, 100)]:
            goals = torch.randint(0, 10, (N, 2), dtype=torch.int32, device="cuda")
            m, mn = timeit(lambda: kernels.darkroom_rollin(goals, 10, H, "uniform", 3, 0, None, 1), R)
            report("darkroom_rollin uniform N=%d H=%d dim=10 (incl. torch.empty of outputs)" % (N, H), N * H, 40, m, mn)
            m, mn = timeit(lambda: kernels.darkroom_rollin(goals, 10, H, "expert", 3, 0, None, 1), R)
            report("darkroom_rollin expert N=%d H=%d dim=10" % (N, H), N * H, 40, m, mn)
    if want("online"):
        arms = torch.tensor(np.random.RandomState(1234).normal(size=(10, 2)) / np.sqrt(2), dtype=torch.float64, device="cuda")   # resident: no H2D copy per pass
        for kind, N, H, d, par in [("opt", 100000, 500, 5, {}), ("emp", 100000, 500, 5, dict(p0=1.0)), ("ucb", 100000, 500, 5, dict(p0=1.0)),
                                   ("thompson", 100000, 500, 5, dict(p0=0.3, p1=0.5, p2=1 / 12.0)),
                                   ("thompson", 100000, 200, 10, dict(p0=0.3, p1=0.0, p2=1.0)),
                                   ("linucb", 100000, 200, 10, dict(p0=1.0, arms=arms)),
                                   ("emp", 10000, 500, 5, dict(p0=1.0)), ("thompson", 10000, 500, 5, dict(p0=0.3, p1=0.5, p2=1 / 12.0))]:
            means, _, _ = kernels.bandit_sample_means(N, d, 0, 0)
            for mat in (True, False):
                m, mn = timeit(lambda: kernels.online_loop(kind, means, H, 0.3, 2, 0, materialise=mat, **par), max(3, R // 4))
                report("online_loop %s N=%d H=%d d=%d materialise=%s" % (kind, N, H, d, mat), N * H, (4 * (3 + d) if mat else 0) + 4, m, mn,
                       trajs_per_s=N / (m * 1e-3))
    if want("gpt2"):
        from dpt_b200.models.net import Transformer
        torch.manual_seed(0)
        for N, H, prec in [(10000, 500, 0), (10000, 500, 1), (2000, 500, 0), (10000, 100, 0), (20000, 500, 1)]:
            m = Transformer({"horizon": H, "state_dim": 1, "action_dim": 5, "n_layer": 4, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
            m.precision = prec
            esz = 2 if prec else 4
            means, _, _ = kernels.bandit_sample_means(N, 5, 0, 0)
            m.online_loop(means, H, 0.3, True, 0, 0)
            mm, mn = timeit(lambda: m.online_loop(means, H, 0.3, True, 1, 0), max(2, R // 4), warm=1)
            kv_bytes = 4 * 2 * 32 * esz * (H * (H - 1) / 2) / H      # K/V bytes read per env-step (mean over t)
            flops = (4 * (24576 * 2) + 832) + 4 * 128 * (H / 2)        # per env-step
            report("gpt2_online_loop %s N=%d H=%d L=4 E=32 sample" % ("bf16-kv" if prec else "fp32", N, H), N * H, kv_bytes + 36, mm, mn,
                   trajs_per_s=N / (mm * 1e-3), kv_cache_gb=N * 4 * 2 * 32 * 512 * esz / 1e9, gflops=N * H * flops / (mm * 1e-3) / 1e9)
            del m
        m = Transformer({"horizon": 100, "state_dim": 2, "action_dim": 5, "n_layer": 4, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
        B, T = 4096, 100
        x = {"query_states": torch.rand(B, 2, device="cuda"), "zeros": torch.zeros(B, 10, device="cuda"), "context_states": torch.rand(B, T, 2, device="cuda"),
             "context_actions": torch.rand(B, T, 5, device="cuda"), "context_next_states": torch.rand(B, T, 2, device="cuda"), "context_rewards": torch.rand(B, T, 1, device="cuda")}
        mm, mn = timeit(lambda: m(x), max(2, R // 4), warm=1)
        report("gpt2_forward fp32 B=%d T=%d (darkroom token layout, test=True)" % (B, T), B * (T + 1), 40, mm, mn, tokens_per_s=B * (T + 1) / (mm * 1e-3))
        m.precision = 1
        mm, mn = timeit(lambda: m(x), max(2, R // 4), warm=1)
        report("gpt2_forward tcgen05 B=%d T=%d (darkroom token layout, test=True)" % (B, T), B * (T + 1), 40, mm, mn, tokens_per_s=B * (T + 1) / (mm * 1e-3))
        del m
        # the bandit model's own context length: 500 tokens, training-style batch (all rows' logits)
        m = Transformer({"horizon": 500, "state_dim": 1, "action_dim": 5, "n_layer": 4, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": False})
        B, T = 1024, 500
        x = {"query_states": torch.ones(B, 1, device="cuda"), "context_states": torch.ones(B, T, 1, device="cuda"),
             "context_actions": torch.rand(B, T, 5, device="cuda"), "context_next_states": torch.ones(B, T, 1, device="cuda"), "context_rewards": torch.rand(B, T, 1, device="cuda")}
        for prec in (0, 1):
            m.precision = prec
            mm, mn = timeit(lambda: m(x), max(2, R // 4), warm=1)
            report("gpt2_forward %s B=%d T=%d (bandit token layout, all rows)" % ("tcgen05" if prec else "fp32", B, T), B * (T + 1), 32, mm, mn, tokens_per_s=B * (T + 1) / (mm * 1e-3))
    if want("gpu_bandit_step"):
        N, d = 100000, 5
        means, _, opt_a = kernels.bandit_sample_means(N, d, 0, 0)
        r = torch.empty(N, device="cuda")
        i = [0]

        def f():
            i[0] += 1
            kernels.gpu_bandit_step(means, opt_a, 0.3, 0, 0, 0, i[0], out=r)
        m, mn = timeit(f, 200)
        report("gpu_bandit_step N=%d d=%d (one launch per env step; latency-bound)" % (N, d), N, 4 * d + 4 + 4, m, mn)


if __name__ == "__main__":
    main()
