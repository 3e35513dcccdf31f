"""Linear-bandit online evaluation (evals/eval_linear_bandit.py): same ``deploy_online_vec``
(:54-97 is a copy of eval_bandit's) with the LinUCB / Thompson controller set of ``online`` (:101-154)."""
from ..ctrls.ctrl_bandit import (BanditTransformerController, EmpMeanPolicy, LinUCBPolicy, OptPolicy,  # noqa: F401
                                 ThompsonSamplingPolicy)
from ..envs.bandit_env import BanditEnvVec, LinearBanditEnv
from .eval_bandit import deploy_online_vec, deploy_online_vec_device, regret_stats  # noqa: F401


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
