"""evals/eval_interactive_bandit.py of the reference: a third copy of ``deploy_online_vec`` (:43-70, identical to
eval_bandit's) and ``run_online_eval`` (:73-142), the cleanest statement of the regret statistics.  Every
controller runs through the fused loops of ``eval_bandit.deploy_online_vec``."""
import numpy as np

from ..ctrls.ctrl_bandit import (BanditTransformerController, EmpMeanPolicy, OptPolicy, ThompsonSamplingPolicy,  # noqa: F401
                                 UCBPolicy)
from ..envs.bandit_env import BanditEnv, BanditEnvVec
from .eval_bandit import deploy_online_vec, deploy_online_vec_device  # noqa: F401


def _sem(v):
    v = np.asarray(v)
    return v.std(axis=0, ddof=1) / np.sqrt(v.shape[0])      # scipy.stats.sem


def generate_eval_trajs(n_eval, dim, bandit_type="uniform"):
    """evals/eval_interactive_bandit.py:29-40: eval bandit instances (means only; no rollout data), drawn on the
    host ``np.random`` stream like the reference."""
    eval_trajs = []
    for _ in range(n_eval):
        if bandit_type == "uniform":
            means = np.random.uniform(0, 1, dim)
        elif bandit_type == "bernoulli":
            means = np.random.beta(1, 1, dim)
        else:
            raise ValueError(f"Unknown bandit_type: {bandit_type}")
        eval_trajs.append({"means": means})
    return eval_trajs


def run_online_eval(eval_trajs, model, n_eval, horizon, var, bandit_type="uniform", sample_model=False):
    """evals/eval_interactive_bandit.py:73-142: Opt, Interactive (transformer), Emp, UCB, Thompson on the same
    tasks; returns the reference's dict (means, sems, regret_means, regret_sems, all_means, all_means_diff)."""
    envs = [BanditEnv(eval_trajs[i]["means"], horizon, var=var, type=bandit_type) for i in range(n_eval)]
    vec_env = BanditEnvVec(envs)
    ctrls = {"opt": OptPolicy(envs, batch_size=len(envs))}
    if model is not None:
        ctrls["Interactive"] = BanditTransformerController(model, sample=sample_model, batch_size=len(envs))
    ctrls["Emp"] = EmpMeanPolicy(envs[0], online=True, batch_size=len(envs))
    ctrls["UCB1.0"] = UCBPolicy(envs[0], const=1.0, batch_size=len(envs))
    ctrls["Thomp"] = ThompsonSamplingPolicy(envs[0], std=var if var > 0 else 0.3, sample=True, prior_mean=0.5,
                                            prior_var=1 / 12.0, warm_start=False, batch_size=len(envs))
    all_means = {}
    for name, c in ctrls.items():
        cm = deploy_online_vec(vec_env, c, horizon).T
        assert cm.shape[0] == n_eval
        all_means[name] = cm
    all_means_diff = {k: all_means["opt"] - v for k, v in all_means.items()}
    cumulative_regret = {k: np.cumsum(v, axis=1) for k, v in all_means_diff.items()}
    return {"means": {k: np.mean(v, axis=0) for k, v in all_means_diff.items()},
            "sems": {k: _sem(v) for k, v in all_means_diff.items()},
            "regret_means": {k: np.mean(v, axis=0) for k, v in cumulative_regret.items()},
            "regret_sems": {k: _sem(v) for k, v in cumulative_regret.items()},
            "all_means": all_means, "all_means_diff": all_means_diff}
