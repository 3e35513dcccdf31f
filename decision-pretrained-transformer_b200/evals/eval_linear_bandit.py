"""Linear-bandit evaluation (evals/eval_linear_bandit.py): same ``deploy_online`` / ``deploy_online_vec``
(:22-97 are copies of eval_bandit's) with the LinUCB / Thompson controller set of ``online`` (:101-154), and the
offline evaluation (:202-330) on the first ``horizon`` rows of each eval trajectory."""
import numpy as np

from ..ctrls.ctrl_bandit import (BanditTransformerController, EmpMeanPolicy, LinUCBPolicy, OptPolicy,  # noqa: F401
                                 ThompsonSamplingPolicy)
from ..envs.bandit_env import BanditEnvVec, LinearBanditEnv
from .eval_bandit import deploy_online, deploy_online_vec, deploy_online_vec_device, regret_stats  # noqa: F401


def online(eval_trajs, model, n_eval, horizon, var):
    """evals/eval_linear_bandit.py:101-170 without plotting."""
    envs = [LinearBanditEnv(eval_trajs[i]["theta"], eval_trajs[i]["arms"], horizon, var=var) for i in range(n_eval)]
    vec_env = BanditEnvVec(envs)
    ctrls = {"opt": OptPolicy(envs, batch_size=len(envs))}
    if model is not None:
        ctrls["Lnr"] = BanditTransformerController(model, sample=True, batch_size=len(envs))
    ctrls["Thomp"] = ThompsonSamplingPolicy(envs[0], std=var, sample=True, prior_mean=0.0, prior_var=1.0,
                                            warm_start=False, batch_size=len(envs))
    ctrls["LinUCB"] = LinUCBPolicy(envs[0], const=1.0, batch_size=len(envs))
    all_means = {name: deploy_online_vec(vec_env, c, horizon).T for name, c in ctrls.items()}
    return all_means, regret_stats(all_means)


def offline(eval_trajs, model, n_eval, horizon, var):
    """evals/eval_linear_bandit.py:202-286 without the bar plot: every controller sees the first ``horizon``
    context rows of each eval trajectory and plays ONE noise-free pull (deploy_eval).  Returns the per-env
    rewards {'opt','lnr','thmp','linreg'} (``lnr`` only when a model is given)."""
    num_envs = len(eval_trajs)
    tmp_env = LinearBanditEnv(eval_trajs[0]["theta"], eval_trajs[0]["arms"], horizon, var=var)
    context_states = np.zeros((num_envs, horizon, tmp_env.dx))
    context_actions = np.zeros((num_envs, horizon, tmp_env.du))
    context_next_states = np.zeros((num_envs, horizon, tmp_env.dx))
    context_rewards = np.zeros((num_envs, horizon, 1))
    envs = []
    for i_eval in range(n_eval):
        traj = eval_trajs[i_eval]
        envs.append(LinearBanditEnv(traj["theta"], traj["arms"], horizon, var=var))
        context_states[i_eval] = traj["context_states"][:horizon]
        context_actions[i_eval] = traj["context_actions"][:horizon]
        context_next_states[i_eval] = traj["context_next_states"][:horizon]
        context_rewards[i_eval] = np.asarray(traj["context_rewards"])[:horizon, None]
    vec_env = BanditEnvVec(envs)
    batch = {"context_states": context_states, "context_actions": context_actions,
             "context_next_states": context_next_states, "context_rewards": context_rewards}
    policies = {"opt": OptPolicy(envs, batch_size=num_envs)}
    if model is not None:
        policies["lnr"] = BanditTransformerController(model, sample=False, batch_size=num_envs)
    policies["thmp"] = ThompsonSamplingPolicy(envs[0], std=var, sample=False, prior_mean=0, prior_var=1.0,
                                              warm_start=False, batch_size=num_envs)
    policies["linreg"] = LinUCBPolicy(envs[0], const=0.0, batch_size=num_envs)
    for name in ("opt", "thmp", "lnr", "linreg"):               # set_batch order of the reference (:263-266)
        if name in policies:
            policies[name].set_batch_numpy_vec(batch)
    baselines = {}
    for name in ("opt", "lnr", "thmp", "linreg"):               # deploy order of the reference (:268-271)
        if name in policies:
            baselines[name] = np.array(vec_env.deploy_eval(policies[name])[3])
    return baselines


def offline_graph(eval_trajs, model, n_eval, horizon, var):
    """evals/eval_linear_bandit.py:289-330 without plotting: suboptimality for every dataset size 1..horizon.
    Returns (horizons, {name: (mean, sem) of opt - name per horizon})."""
    horizons = np.linspace(1, horizon, horizon, dtype=int)
    sub_mean, sub_sem = {}, {}
    for h in horizons:
        b = offline(eval_trajs, model, n_eval=n_eval, horizon=int(h), var=var)
        for k, v in b.items():
            if k == "opt":
                continue
            d = b["opt"] - v
            sub_mean.setdefault(k, []).append(np.mean(d))
            sub_sem.setdefault(k, []).append(np.std(d, ddof=1) / np.sqrt(len(d)) if len(d) > 1 else np.nan)   # scipy.stats.sem
    return horizons, {k: (np.array(sub_mean[k]), np.array(sub_sem[k])) for k in sub_mean}
