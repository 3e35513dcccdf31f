#!/usr/bin/env python
"""bench.py -- rollout-collection throughput of the DPT hot path on B200.

Workload (BASELINE.json configs[4] per-GPU shard, the configuration the env-steps/s target is quoted
on): bandit rollin_bandit collection, H=500, dim=5, var=0.3, 125 000 envs per GPU (weak scaling:
1M envs on 8 GPUs), Philox noise, outputs in the reference consumer's fp32 layout (32 B per
env-step, 2 GB per step per GPU -- larger than the 126 MB L2, so no flush is needed between steps).
A "step" is one pass of the fused kernel over the GPU's env shard.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, DIM, VAR = 500, 5, 0.3
ENVS_PER_GPU = int(os.environ.get("DPT_BENCH_ENVS", 125000))
BYTES_PER_STEP = 4 * (2 * 1 + DIM + 1)          # four fp32 context rows per env-step
METRIC = "env-steps/sec (bandit rollout collect, H=500 dim=5)"
UNIT = "env-steps/s"


def workload_name(n_gpus):
    return ("bandit rollin_bandit collection H=%d dim=%d var=%.1f, %d envs/GPU x %d GPU (BASELINE configs[4] shard), "
            "outputs %.2f GB/step/GPU > L2 (no flush needed)" % (H, DIM, VAR, ENVS_PER_GPU, n_gpus,
                                                                  ENVS_PER_GPU * H * BYTES_PER_STEP / 1e9))


# ----------------------------------------------------------------------------- clocks ---------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU via NVML while the timed region runs."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
            0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, uuid):
        self.samples, self.reasons, self.max_mhz, self._stop, self.ok = [], set(), None, threading.Event(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:   # noqa: BLE001
            self.err = repr(e)

    def _once(self):
        nv = self.nv
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            m = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            m = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in self.BITS.items():
            if m & bit:
                self.reasons.add(name)

    def _loop(self):
        while not self._stop.is_set():
            self._once()
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self._once()
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self):
        if self.ok:
            self._stop.set()
            self.t.join()
            self._once()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- reference arm --
def reference_sample(pool, envs_per_core):
    steps, wall = pool.run(envs_per_core, DIM, H, VAR)
    return steps, wall


def run_reference(args, rank):
    """The reference's own CPU implementation of the path (its oracle port: the reference is pure
    Python and cannot travel to the GPU box) on all host cores, one process per core."""
    if rank != 0:
        return
    from oracle import cpu_bench
    pool = cpu_bench.BanditRollinPool()
    envs_per_core = int(os.environ.get("DPT_REF_ENVS_PER_CORE", 100))
    for _ in range(args.warmup):
        pool.run(envs_per_core, DIM, H, VAR)
    tot_steps, t0 = 0, time.perf_counter()
    for i in range(args.steps):
        s, _ = pool.run(envs_per_core, DIM, H, VAR, seed0=1000 * (i + 1))
        tot_steps += s
    wall = time.perf_counter() - t0
    pool.close()
    v = tot_steps / wall
    sample = "%d cores x %d envs x H=%d per step (oracle port of collect_data.generate_bandit_histories)" % (
        pool.cores, envs_per_core, H)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": pool.cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def other_workloads(kernels, torch):
    """The other BASELINE.json configs on one GPU (informational; CUDA events, 3 warm-up + 5 timed runs each)."""
    import numpy as np

    def t(fn, reps=5, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    out = {}
    goals = torch.tensor(np.repeat(np.stack(np.meshgrid(np.arange(10), np.arange(10), indexing="ij"), -1).reshape(-1, 2), 1000, 0),
                         dtype=torch.int32, device="cuda")
    ms = t(lambda: kernels.darkroom_rollin(goals, 10, 100, "uniform", 1, 0, None, 1))
    out["config2_darkroom_rollin_100k_envs_H100"] = {"ms": ms, "env_steps_per_s": 1e7 / (ms * 1e-3), "gbs": 4e8 / (ms * 1e-3) / 1e9}
    means10, _, _ = kernels.bandit_sample_means(100000, 10, 0, 0)
    arms = np.random.RandomState(1234).normal(size=(10, 2)) / np.sqrt(2)
    ms = t(lambda: kernels.online_loop("thompson", means10, 200, 0.3, 1, 0, p0=0.3, p1=0.0, p2=1.0))
    out["config3_linear_thompson_collect_100k_envs_H200_d10"] = {"ms": ms, "env_steps_per_s": 2e7 / (ms * 1e-3), "trajs_per_s": 1e5 / (ms * 1e-3)}
    ms = t(lambda: kernels.online_loop("linucb", means10, 200, 0.3, 1, 0, p0=1.0, arms=arms, materialise=False))
    out["config3_linucb_online_100k_envs_H200_d10"] = {"ms": ms, "env_steps_per_s": 2e7 / (ms * 1e-3), "trajs_per_s": 1e5 / (ms * 1e-3)}
    import dpt_b200
    from dpt_b200.models.net import Transformer
    torch.manual_seed(0)
    m = Transformer({"horizon": 500, "state_dim": 1, "action_dim": 5, "n_layer": 4, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
    means5, _, _ = kernels.bandit_sample_means(10000, 5, 0, 0)
    for prec, name in ((0, "fp32"), (1, "bf16_kv")):
        m.precision = prec
        ms = t(lambda: m.online_loop(means5, 500, 0.3, True, 1, 0), reps=2, warm=1)
        esz = 2 if prec else 4
        out["config4_gpt2_online_eval_10k_envs_H500_" + name] = {
            "ms": ms, "trajs_per_s": 1e4 / (ms * 1e-3), "env_steps_per_s": 5e6 / (ms * 1e-3),
            "kv_read_gbs": 1e4 * 4 * 2 * 32 * esz * (500 * 499 / 2) / (ms * 1e-3) / 1e9}
    return out


# ----------------------------------------------------------------------------- our arm --------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-other", action="store_true", help="skip the secondary workloads (configs 2-4)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)   # timing rule: at least 3 warm-up steps (the JSON line reports the value used)
    args.steps = max(args.steps, 1)

    # CPU baseline first (rank 0, N=1 only), before the GPU is busy: bounded sample on all host cores
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_bench
        pool = cpu_bench.BanditRollinPool()
        epc = int(os.environ.get("DPT_CPU_ENVS_PER_CORE", 3000))
        steps, wall = pool.run(epc, DIM, H, VAR)
        pool.close()
        cpu_baseline = {"value": steps / wall, "unit": UNIT, "cores": pool.cores, "kind": "port",
                        "sample": "%d cores x %d envs x H=%d, oracle port of collect_data.generate_bandit_histories "
                                  "(%.2f s wall, %.2f s mean per worker)" % (pool.cores, epc, H, wall,
                                                                             sum(pool.last_worker_seconds) / pool.cores)}

    import torch
    import torch.distributed as dist
    import dpt_b200
    from dpt_b200 import kernels
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = ENVS_PER_GPU
    env_id0 = rank * N                        # contiguous global env ranges
    seed = 0
    means, _, _ = kernels.bandit_sample_means(N, DIM, seed, env_id0)
    out = {"context_states": torch.empty((N, H, 1), device=dev), "context_actions": torch.empty((N, H, DIM), device=dev),
           "context_next_states": torch.empty((N, H, 1), device=dev), "context_rewards": torch.empty((N, H, 1), device=dev)}
    # per-step return statistics: one pre-zeroed slot per step, so the timed loop launches nothing but the
    # fused kernel (and, multi-GPU, the async NCCL gather of that step's [3] statistics, which overlaps the
    # next step's kernel)
    n_slots = args.steps + args.warmup
    stats = torch.zeros((n_slots, 3), dtype=torch.float64, device=dev)
    gathered = torch.zeros((n_slots, 3 * world), dtype=torch.float64, device=dev)
    pending = []
    peer, gather_how = None, "none"
    if world > 1:
        from dpt_b200 import dist as ddist
        try:      # all-gather fused into the kernel: the last CTA stores the totals into every rank's buffer over NVLink
            if os.environ.get("DPT_BENCH_GATHER", "p2p") != "p2p":
                raise RuntimeError("NCCL gather requested")
            peer = ddist.PeerGather(n_slots)
            gather_how = "fused in-kernel all-gather of return stats over NVLink peer memory (CUDA IPC), no collective launch"
        except Exception as e:   # noqa: BLE001
            peer = None
            gather_how = "NCCL all-gather of return stats every step (async) [p2p unavailable: %s]" % str(e)[:80]

    def step(i):
        if peer is not None:
            kernels.bandit_rollin(means, H, VAR, seed + i, env_id0, out=out, stats=stats[i], peer=peer, peer_slot=i)
        else:
            kernels.bandit_rollin(means, H, VAR, seed + i, env_id0, out=out, stats=stats[i])
            if world > 1:   # every pass's statistics are gathered; the call is asynchronous (overlaps the next pass)
                pending.append(dist.all_gather_into_tensor(gathered[i], stats[i], async_op=True))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler("GPU-" + str(torch.cuda.get_device_properties(dev).uuid)) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    if sampler:
        sampler.start()
    ev[0].record()
    for i in range(args.steps):
        step(args.warmup + i)
        ev[i + 1].record()
    for w_ in pending:      # the statistics exchange belongs to the timed region
        w_.wait()
    end_ev = torch.cuda.Event(enable_timing=True)
    end_ev.record()
    barrier()
    if sampler:
        sampler.stop()
    total_ms = ev[0].elapsed_time(end_ev)
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t[0])
    value = world * N * H * args.steps / (total_ms * 1e-3)

    # roofline of the dominant kernel (bandit_rollin_fast<5>): algorithmic bytes / mean launch duration.
    # A step IS one launch of it, so the per-step event deltas are its launch durations.
    kern_ms = statistics.mean(per)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = N * H * BYTES_PER_STEP / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "bandit_rollin_fast<5,PHILOX>", "algorithmic_bytes_per_launch": N * H * BYTES_PER_STEP,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "launch_ms_mean": kern_ms, "launch_ms_min": min(per)}
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tfile):
        try:
            roofline["traffic"] = json.load(open(tfile)).get("bandit_rollin_fast_dram_bytes_per_launch")
        except Exception:
            pass

    # e2e: the same collection through the host-buffer C-ABI call (means in pinned host memory,
    # contexts returned to pinned host memory), copies inside the timed region.
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 10))
    means_host = means.cpu().pin_memory()
    host_out, scratch = kernels.bandit_rollin_host(means_host, H, VAR, seed, env_id0)
    for i in range(2):
        kernels.bandit_rollin_host(means_host, H, VAR, seed + i, env_id0, out=host_out, scratch=scratch)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        kernels.bandit_rollin_host(means_host, H, VAR, seed + 100 + i, env_id0, out=host_out, scratch=scratch)
    e1.record()
    d2h_last = int(dpt_b200._lib.lib().dpt_bandit_rollin_host_last_d2h_bytes())
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0])
    e2e = {"value": world * N * H * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": N * DIM * 4,
           # what crossed PCIe in the last timed call: chunks returned by DMA carry actions + rewards (4*(d+1) B per
           # step), chunks returned in compact form 5 B per step (expanded by host threads); the constant state
           # columns are always written on the host
           "d2h_bytes_per_step": d2h_last, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
           "api": "dpt_bandit_rollin_host (pinned host buffers, chunked H2D/kernel/D2H pipeline, hybrid DMA / host expansion)"}

    for w_ in pending:
        w_.wait()
    torch.cuda.synchronize()
    if peer is not None:
        barrier()
        totals = torch.tensor(peer.read().sum((0, 1)))
        peer.close()
    elif world > 1:
        totals = gathered.view(-1, 3).sum(0)
    else:
        totals = stats.sum(0)
    other = None
    if rank == 0 and world == 1 and not args.no_other:
        other = other_workloads(kernels, torch)
    if rank == 0:
        st = totals
        n_tot = world * N * H * (args.steps + args.warmup)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(world), "envs_per_gpu": N, "H": H, "dim": DIM, "var": VAR,
                       "noise": "philox4x32-10", "l2": "outputs larger than L2, no flush",
                       "parallelism": "env-sharded x%d%s" % (world, ", " + gather_how if world > 1 else "")},
            "roofline": roofline, "e2e": e2e, "gpu_launches": args.steps,
            "clocks": sampler.summary() if sampler else None,
            "return_stats": {"mean_reward": float(st[0]) / n_tot, "frac_optimal_arm": float(st[2]) / n_tot},
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if other is not None:
            line["other_workloads"] = other
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
