"""CPU baseline legs for bench.py -- TEST INFRASTRUCTURE (the checker timed as a baseline, never
the product).  Each worker runs the oracle port of the reference's rollout on one host core; the
pool uses every core the process may run on, because the reference itself is single-threaded
Python and the only way to use a whole host with it is one process per core.
"""
import os
import time


def _bandit_worker(args):
    seed, n_envs, dim, H, var = args
    import numpy as np
    from oracle import dpt_oracle as O
    np.random.seed(seed)
    t0 = time.perf_counter()
    trajs = O.generate_bandit_histories(n_envs, dim, H, var, O.GlobalNoise(record=False))
    dt = time.perf_counter() - t0
    chk = float(sum(t["context_rewards"].sum() for t in trajs))
    return n_envs * H, dt, chk


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class BanditRollinPool:
    """Persistent spawn pool (safe next to an initialised CUDA context)."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        self.cores = cores or host_cores()
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self.pool.map(_bandit_worker, [(i, 1, 5, 8, 0.3) for i in range(self.cores)])   # start-up + imports

    def run(self, envs_per_core, dim, H, var, seed0=0):
        """One sample: every core rolls ``envs_per_core`` envs.  Returns (env_steps, wall_seconds)."""
        t0 = time.perf_counter()
        res = self.pool.map(_bandit_worker, [(seed0 + i, envs_per_core, dim, H, var) for i in range(self.cores)])
        wall = time.perf_counter() - t0
        self.last_worker_seconds = [r[1] for r in res]
        return sum(r[0] for r in res), wall

    def close(self):
        self.pool.close()
        self.pool.join()
