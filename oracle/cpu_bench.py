"""CPU baseline legs for bench.py -- TEST INFRASTRUCTURE (the checker timed as a baseline, never
the product).

Two kinds of worker, both one process per host core (the reference is single-threaded Python, so
the only way to use a whole host with it is one process per core -- BASELINE.md §4):

* ``kind = "reference"``: the UNMODIFIED reference's own functions, imported from ``baseline/_ref``
  (staged by ``oracle.ref_loader.stage()``; ``/root/reference`` in the dev container) behind the four
  import shims of ``oracle/ref_loader.py``:
    bandit      collect_data.generate_bandit_histories            (collect_data.py:221-225)      config 1 / 5
    darkroom    collect_data.generate_darkroom_histories          (collect_data.py:290-293)      config 2
    lin_thomp   collect_data.rollin_linear_bandit_vec             (collect_data.py:56-80)        config 3a
    lin_ucb     evals.eval_linear_bandit.deploy_online_vec+LinUCB (eval_linear_bandit.py:54-97)  config 3b
    gpt2_online evals.eval_bandit.deploy_online_vec + BanditTransformerController
                                                                  (eval_bandit.py:56-103, ctrl_bandit.py:383-444) config 4
    emp/ucb/thompson online: same loop with the classical controllers (ctrl_bandit.py:91,351,232)
    darkroom_online evals.eval_darkroom.deploy_online_vec + DarkroomTransformerController (eval_darkroom.py:20-84)    §8 (f)1
* ``kind = "port"``: the numpy restatement in ``oracle/dpt_oracle.py`` (bit-identical outputs from the
  same np.random stream; faster per core than the reference because it builds the categorical cdf once
  per env instead of letting np.random.choice re-validate ``p`` at every draw).
"""
import os
import time


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _ref():
    from oracle import ref_loader
    return ref_loader.load()


def _one_thread():
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:   # noqa: BLE001
        pass


def _work(args):
    """(kind, workload, seed, n_envs, H, extra) -> (env_steps, trajectories, seconds, checksum)"""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):   # the reference prints progress lines; bench.py prints ONE JSON line
        return _work_impl(args)


def _work_impl(args):
    kind, workload, seed, n, H, extra = args
    import numpy as np
    np.random.seed(seed)
    if kind == "port":
        from oracle import dpt_oracle as O
        assert workload == "bandit"
        t0 = time.perf_counter()
        trajs = O.generate_bandit_histories(n, extra["dim"], H, extra["var"], O.GlobalNoise(record=False))
        dt = time.perf_counter() - t0
        return n * H, n, dt, float(sum(t["context_rewards"].sum() for t in trajs))
    ref = _ref()
    if extra.get("threads"):
        import torch
        torch.set_num_threads(int(extra["threads"]))
    else:
        _one_thread()
    var = extra.get("var", 0.3)
    if workload == "bandit":
        t0 = time.perf_counter()
        trajs = ref.collect_data.generate_bandit_histories(n, extra["dim"], H, var, n_hists=1, n_samples=1, cov=0.0, type="uniform")
        dt = time.perf_counter() - t0
        return n * H, n, dt, float(sum(t["context_rewards"].sum() for t in trajs))
    if workload == "darkroom":
        dim = extra["dim"]
        goals = [(i % dim, (i // dim) % dim) for i in range(n)]
        t0 = time.perf_counter()
        trajs = ref.collect_data.generate_darkroom_histories(goals, dim, H, n_hists=1, n_samples=1, rollin_type="uniform")
        dt = time.perf_counter() - t0
        return n * H, n, dt, float(sum(t["context_rewards"].sum() for t in trajs))
    if workload in ("lin_thomp", "lin_ucb"):
        dim, lin_d = extra["dim"], extra["lin_d"]
        arms = np.random.RandomState(seed=1234).normal(size=(dim, lin_d)) / np.sqrt(lin_d)   # collect_data.py:230-231
        envs = [ref.bandit_env.sample_linear(arms, H, var) for _ in range(n)]
        t0 = time.perf_counter()
        if workload == "lin_thomp":
            out = ref.collect_data.rollin_linear_bandit_vec(envs)
            chk = float(out[3].sum())
        else:
            vec = ref.bandit_env.BanditEnvVec(envs)
            cum = ref.eval_linear_bandit.deploy_online_vec(vec, ref.ctrl_bandit.LinUCBPolicy(envs[0], const=1.0, batch_size=n), H)
            chk = float(np.sum(cum))
        dt = time.perf_counter() - t0
        return n * H, n, dt, chk
    if workload == "darkroom_online":   # evals/eval_darkroom.py:20-84 + ctrls/ctrl_darkroom.py:23-66
        import torch
        dim, horizon, Heps = extra["dim"], extra["horizon"], extra["Heps"]
        torch.manual_seed(0)
        cfg = {"horizon": H, "state_dim": 2, "action_dim": 5, "n_layer": extra.get("n_layer", 4), "n_embd": 32, "n_head": 1,
               "dropout": 0.0, "test": True}
        model = ref.net.Transformer(cfg).to(ref.net.device).eval()
        envs = [ref.darkroom_env.DarkroomEnv(dim, ((seed + i) % dim, (3 * i + 1) % dim), horizon) for i in range(n)]
        ctrl = ref.ctrl_darkroom.DarkroomTransformerController(model, batch_size=n, sample=True)
        t0 = time.perf_counter()
        ret = ref.eval_darkroom.deploy_online_vec(ref.darkroom_env.DarkroomEnvVec(envs), ctrl, Heps, H, horizon)
        dt = time.perf_counter() - t0
        return n * Heps * horizon, n, dt, float(np.sum(ret))
    # online loops on the plain bandit (config 4 and its classical controllers)
    d = extra["dim"]
    means = np.random.uniform(0, 1, (n, d))
    envs = [ref.bandit_env.BanditEnv(m, H, var=var) for m in means]
    vec = ref.bandit_env.BanditEnvVec(envs)
    C = ref.ctrl_bandit
    if workload == "gpt2_online":
        import torch
        torch.manual_seed(0)
        cfg = {"horizon": H if not extra.get("model_H") else extra["model_H"], "state_dim": 1, "action_dim": d,
               "n_layer": extra.get("n_layer", 4), "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True}
        model = ref.net.Transformer(cfg).to(ref.net.device).eval()
        ctrl = C.BanditTransformerController(model, sample=True, batch_size=n)
    elif workload == "emp":
        ctrl = C.EmpMeanPolicy(envs[0], online=True, batch_size=n)
    elif workload == "ucb":
        ctrl = C.UCBPolicy(envs[0], const=1.0, batch_size=n)      # runs only at n == 200 (ctrl_bandit.py:374)
    elif workload == "thompson":
        ctrl = C.ThompsonSamplingPolicy(envs[0], std=var, sample=True, prior_mean=0.5, prior_var=1 / 12.0, warm_start=False, batch_size=n)
    elif workload == "opt":
        ctrl = C.OptPolicy(envs, batch_size=n)
    else:
        raise ValueError(workload)
    t0 = time.perf_counter()
    cum = ref.eval_bandit.deploy_online_vec(vec, ctrl, H)
    dt = time.perf_counter() - t0
    return n * H, n, dt, float(np.sum(cum))


def _cpu_only():
    """Pool initializer: the workers are the CPU arm.  The reference picks ``torch.device('cuda' if available)`` at import
    (ctrls/ctrl_bandit.py:8, models/net.py), so on a GPU box its transformer controllers would otherwise run their forwards on
    the GPU from 16 processes -- hide the device before anything imports torch in the worker."""
    os.environ["CUDA_VISIBLE_DEVICES"] = ""


class CpuPool:
    """Persistent spawn pool (safe next to an initialised CUDA context), one worker per host core, no GPU visible."""

    def __init__(self, cores=None, kind=None):
        import multiprocessing as mp
        from oracle import ref_loader
        self.cores = cores or host_cores()
        self.kind = kind or ("reference" if ref_loader.available() else "port")
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_only)
        self.pool.map(_work, [(self.kind, "bandit", i, 1, 8, {"dim": 5, "var": 0.3}) for i in range(self.cores)])   # start-up + imports

    def run(self, workload, n_per_core, H, seed0=0, kind=None, **extra):
        """One sample: every core runs ``n_per_core`` envs.  Returns dict(env_steps, trajs, wall, worker_s)."""
        kind = kind or self.kind
        t0 = time.perf_counter()
        res = self.pool.map(_work, [(kind, workload, seed0 + i, n_per_core, H, extra) for i in range(self.cores)])
        wall = time.perf_counter() - t0
        return {"env_steps": sum(r[0] for r in res), "trajs": sum(r[1] for r in res), "wall": wall,
                "worker_s": sum(r[2] for r in res) / len(res), "cores": self.cores, "kind": kind}

    def close(self):
        self.pool.close()
        self.pool.join()


class BanditRollinPool(CpuPool):   # round-1 name, kept for scripts: the oracle-port collection leg
    def __init__(self, cores=None):
        super().__init__(cores, kind="port")

    def run(self, envs_per_core, dim, H, var, seed0=0):   # noqa: D102
        r = CpuPool.run(self, "bandit", envs_per_core, H, seed0, dim=dim, var=var)
        self.last_worker_seconds = [r["worker_s"]] * self.cores
        return r["env_steps"], r["wall"]
