"""Online in-context evaluation with the reference's interface (evals/eval_bandit.py).

``deploy_online_vec(vec_env, controller, horizon, include_meta=False)`` keeps its signature and
return values (cum_means [H,N] float64, meta = four [N,H,.] float64 arrays).  When the env is this
package's ``BanditEnvVec`` and the controller exposes ``fused_spec()`` the whole loop -- controller
statistics, arm choice, env step, context rows, expected reward of the chosen arm, per-step regret
sums -- runs in ONE fused launch (dpt_online_loop / dpt_gpt2_online_loop) instead of H Python
iterations over N env objects.  Anything else takes the generic loop below, which is the reference's
loop verbatim in structure and drives ``vec_env.deploy`` one step at a time.
"""
import numpy as np
import torch

from .. import kernels, rng
from ..ctrls.ctrl_bandit import (BanditTransformerController, EmpMeanPolicy, GreedyOptPolicy, OptPolicy,  # noqa: F401
                                 PessMeanPolicy, ThompsonSamplingPolicy, UCBPolicy)
from ..envs.bandit_env import BanditEnv, BanditEnvVec


def _fusable(vec_env, controller):
    if not isinstance(vec_env, BanditEnvVec) or not callable(getattr(controller, "fused_spec", None)):
        return None
    if len({float(e.var) for e in vec_env.envs}) != 1 or len({e.type for e in vec_env.envs}) != 1:
        return None
    return controller.fused_spec()


def deploy_online_vec_device(vec_env, controller, horizon, include_meta=False, regret=True, seed=None, env_id0=0,
                             inject=None, dump=False):
    """Fused loop, results left on the device (dict of torch tensors, see kernels.online_loop)."""
    spec = _fusable(vec_env, controller)
    if spec is None:
        raise NotImplementedError("controller / env pair has no fused path")
    key = rng.next_key() if seed is None else seed
    var, rtype = float(vec_env.envs[0].var), vec_env.envs[0].type
    if spec["kind"] == "transformer":
        return controller.fused_online_loop(vec_env.means, horizon, var, key, env_id0, include_meta, regret, inject, dump,
                                            rtype)
    return kernels.online_loop(spec["kind"], vec_env.means, horizon, var, key, env_id0, spec.get("p0", 0.0),
                               spec.get("p1", 0.0), spec.get("p2", 0.0), spec.get("arms"), include_meta, regret,
                               inject, dump, rtype)


def deploy_online(env, controller, horizon):
    """evals/eval_bandit.py:24-53 (= eval_linear_bandit.py:22-51): the single-env loop with torch context
    tensors; the controller sees the first h rows through ``set_batch``.  Returns cum_means [horizon]."""
    dev = kernels._dev()
    f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).to(dev)   # noqa: E731
    context_states = torch.zeros((1, horizon, env.dx), device=dev)
    context_actions = torch.zeros((1, horizon, env.du), device=dev)
    context_next_states = torch.zeros((1, horizon, env.dx), device=dev)
    context_rewards = torch.zeros((1, horizon, 1), device=dev)
    cum_means = []
    for h in range(horizon):
        controller.set_batch({"context_states": context_states[:, :h, :], "context_actions": context_actions[:, :h, :],
                              "context_next_states": context_next_states[:, :h, :],
                              "context_rewards": context_rewards[:, :h, :]})
        states_lnr, actions_lnr, next_states_lnr, rewards_lnr = env.deploy(controller)
        context_states[0, h, :] = f(states_lnr[0])
        context_actions[0, h, :] = f(actions_lnr[0])
        context_next_states[0, h, :] = f(next_states_lnr[0])
        context_rewards[0, h, :] = f(rewards_lnr[0])
        cum_means.append(env.get_arm_value(np.asarray(actions_lnr).flatten()))
    return np.array(cum_means)


def deploy_online_vec(vec_env, controller, horizon, include_meta=False):
    """evals/eval_bandit.py:56-103."""
    if _fusable(vec_env, controller) is not None:
        out = deploy_online_vec_device(vec_env, controller, horizon, include_meta, regret=False)
        cum_means = out["cum_means"].cpu().numpy().astype(np.float64)
        if not include_meta:
            return cum_means
        meta = {k: out[k].cpu().numpy().astype(np.float64) for k in
                ("context_states", "context_actions", "context_next_states", "context_rewards")}
        return cum_means, meta

    num_envs = vec_env.num_envs
    context_states = np.zeros((num_envs, horizon, vec_env.dx))
    context_actions = np.zeros((num_envs, horizon, vec_env.du))
    context_next_states = np.zeros((num_envs, horizon, vec_env.dx))
    context_rewards = np.zeros((num_envs, horizon, 1))
    cum_means = []
    for h in range(horizon):
        batch = {"context_states": context_states[:, :h, :], "context_actions": context_actions[:, :h, :],
                 "context_next_states": context_next_states[:, :h, :], "context_rewards": context_rewards[:, :h, :]}
        controller.set_batch_numpy_vec(batch)
        states_lnr, actions_lnr, next_states_lnr, rewards_lnr = vec_env.deploy(controller)
        context_states[:, h, :] = states_lnr
        context_actions[:, h, :] = actions_lnr
        context_next_states[:, h, :] = next_states_lnr
        context_rewards[:, h, :] = rewards_lnr[:, None]
        cum_means.append(vec_env.get_arm_value(actions_lnr))
    cum_means = np.array(cum_means)
    if not include_meta:
        return cum_means
    return cum_means, {"context_states": context_states, "context_actions": context_actions,
                       "context_next_states": context_next_states, "context_rewards": context_rewards}


def regret_stats(all_means):
    """evals/eval_bandit.py:169-178: dict name -> [N,H] expected reward -> per-step and cumulative
    regret mean / standard error (scipy.stats.sem, ddof=1) against all_means['opt']."""
    opt = np.asarray(all_means["opt"])
    n = opt.shape[0]
    out = {}
    for k, v in all_means.items():
        diff = opt - np.asarray(v)
        cr = np.cumsum(diff, axis=1)
        out[k] = {"mean": diff.mean(0), "sem": diff.std(0, ddof=1) / np.sqrt(n),
                  "regret_mean": cr.mean(0), "regret_sem": cr.std(0, ddof=1) / np.sqrt(n)}
    return out


def online(eval_trajs, model, n_eval, horizon, var, bandit_type="uniform"):
    """evals/eval_bandit.py:107-178 without the matplotlib part: runs Opt, the transformer (if a
    model is given), EmpMean(online), UCB(1.0) and Thompson on the same tasks and returns
    (all_means dict of [n_eval,H], regret statistics)."""
    envs = [BanditEnv(eval_trajs[i]["means"], horizon, var=var) for i in range(n_eval)]
    vec_env = BanditEnvVec(envs)
    ctrls = {"opt": OptPolicy(envs, batch_size=len(envs))}
    if model is not None:
        ctrls["Lnr"] = BanditTransformerController(model, sample=True, batch_size=len(envs))
    ctrls["Emp"] = EmpMeanPolicy(envs[0], online=True, batch_size=len(envs))
    ctrls["UCB1.0"] = UCBPolicy(envs[0], const=1.0, batch_size=len(envs))
    ctrls["Thomp"] = ThompsonSamplingPolicy(envs[0], std=var, sample=True, prior_mean=0.5, prior_var=1 / 12.0,
                                            warm_start=False, batch_size=len(envs))
    all_means = {}
    for name, c in ctrls.items():
        cm = deploy_online_vec(vec_env, c, horizon).T
        assert cm.shape[0] == n_eval
        all_means[name] = cm
    return all_means, regret_stats(all_means)


def offline(eval_trajs, model, n_eval, horizon, var, bandit_type="uniform", np_random_compat=False):
    """evals/eval_bandit.py:214-301 without the bar plot: every controller sees the first ``horizon``
    context rows of each eval trajectory and plays ONE noise-free pull (deploy_eval); returns the dict
    of per-env rewards {'opt','lnr','emp','thmp','lcb'} (``lnr`` only when a model is given).  Per-arm
    statistics come from dpt_arm_stats, the transformer decision from dpt_gpt2_forward."""
    num_envs = len(eval_trajs)
    tmp_env = BanditEnv(eval_trajs[0]["means"], horizon, var=var)
    context_states = np.zeros((num_envs, horizon, tmp_env.dx))
    context_actions = np.zeros((num_envs, horizon, tmp_env.du))
    context_next_states = np.zeros((num_envs, horizon, tmp_env.dx))
    context_rewards = np.zeros((num_envs, horizon, 1))
    envs = []
    for i_eval in range(n_eval):
        traj = eval_trajs[i_eval]
        envs.append(BanditEnv(traj["means"], horizon, var=var))
        context_states[i_eval] = traj["context_states"][:horizon]
        context_actions[i_eval] = traj["context_actions"][:horizon]
        context_next_states[i_eval] = traj["context_next_states"][:horizon]
        context_rewards[i_eval] = np.asarray(traj["context_rewards"])[:horizon, None]
    vec_env = BanditEnvVec(envs)
    vec_env.np_random_compat = np_random_compat     # keep np.random aligned with the reference's env draws
    batch = {"context_states": context_states, "context_actions": context_actions,
             "context_next_states": context_next_states, "context_rewards": context_rewards}
    policies = {"opt": OptPolicy(envs, batch_size=num_envs),
                "emp": EmpMeanPolicy(envs[0], online=False, batch_size=num_envs)}
    if model is not None:
        policies["lnr"] = BanditTransformerController(model, sample=False, batch_size=num_envs)
    policies["lcb"] = PessMeanPolicy(envs[0], const=.8, batch_size=len(envs))
    policies["thmp"] = ThompsonSamplingPolicy(envs[0], std=var, sample=False, prior_mean=0.5, prior_var=1 / 12.0,
                                              warm_start=False, batch_size=num_envs)
    for name in ("opt", "emp", "thmp", "lcb", "lnr"):            # set_batch order of the reference (:270-274)
        if name in policies:
            policies[name].set_batch_numpy_vec(batch)
    baselines = {}
    for name in ("opt", "emp", "lnr", "lcb", "thmp"):            # deploy order of the reference (:276-280)
        if name in policies:
            baselines[name] = np.array(vec_env.deploy_eval(policies[name])[3])
    return baselines


def offline_graph(eval_trajs, model, n_eval, horizon, var, bandit_type="uniform"):
    """evals/eval_bandit.py:304-335 without plotting: suboptimality against dataset size.  Returns
    (horizons [50], {name: regret-of-the-mean per horizon})."""
    horizons = np.linspace(1, horizon, 50, dtype=int)
    all_means = []
    for h in horizons:
        b = offline(eval_trajs, model, n_eval=n_eval, horizon=int(h), var=var, bandit_type=bandit_type)
        all_means.append({k: np.mean(v, axis=0) for k, v in b.items()})
    regrets = {k: np.array([m["opt"] - m[k] for m in all_means]) for k in all_means[0] if k != "opt"}
    return horizons, regrets
