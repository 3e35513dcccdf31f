"""Rollout collection with the reference's interface (collect_data.py), backed by the CUDA path.

Reference functions kept drop-in: ``rollin_bandit(env, cov, orig=False)``, ``rollin_mdp(env,
rollin_type)``, ``rollin_linear_bandit_vec(envs)``, ``generate_*_histories*`` (same traj-dict keys,
dtypes and shapes as collect_data.py:158-300).  Where the reference loops over env objects in
Python, the ``generate_*`` functions here make ONE fused launch over all envs; the device-resident
batch interface (``collect_bandit`` / ``collect_darkroom``) skips the host format entirely.
"""
import numpy as np
import torch

from . import kernels, rng


# --------------------------------------------------------------------------- device batches ---
def collect_bandit(n_envs, dim, horizon, var, seed=None, env_id0=0, means=None, device=None, reward_type="uniform"):
    """Task draw + rollin_bandit for ``n_envs`` envs, all on the device.  Returns a dict of fp32
    device tensors: means [N,d], opt_a_index [N], optimal_actions [N,d], context_* [N,H,.]."""
    key = rng.next_key() if seed is None else seed
    if means is None:
        means, opt_idx, opt_a = kernels.bandit_sample_means(n_envs, dim, key, env_id0, device)
    else:
        means = kernels._as(means, torch.float32, kernels._dev(device))
        opt_idx, opt_a = kernels.bandit_opt_action(means)
    out = kernels.bandit_rollin(means, horizon, float(var), key, env_id0, reward_type=reward_type)
    out.update(means=means, opt_a_index=opt_idx, optimal_actions=opt_a)
    return out


def collect_darkroom(goals, dim, horizon, rollin_type="uniform", perm_index=None, n_samples=1, seed=None, env_id0=0):
    """rollin_mdp + query states / optimal actions for all envs in one launch (device tensors)."""
    key = rng.next_key() if seed is None else seed
    return kernels.darkroom_rollin(goals, dim, horizon, rollin_type, key, env_id0, perm_index, n_samples)


# --------------------------------------------------------------------------- single-env API ---
def rollin_bandit(env, cov, orig=False):
    """collect_data.py:23-53.  ``cov`` is ignored exactly like the reference (overwritten at :30).
    Returns xs (H,1) int64, us (H,d) f64, xps (H,1) int64, rs (H,) f64."""
    H = env.H_context
    out = kernels.bandit_rollin(torch.as_tensor(np.asarray(env.means)[None, :], dtype=torch.float32), H,
                                float(env.var), rng.next_key(), 0, reward_type=getattr(env, "type", "uniform"))
    xs = out["context_states"][0].cpu().numpy().astype(np.int64)
    us = out["context_actions"][0].cpu().numpy().astype(np.float64)
    xps = out["context_next_states"][0].cpu().numpy().astype(np.int64)
    rs = out["context_rewards"][0, :, 0].cpu().numpy().astype(np.float64)
    return xs, us, xps, rs


def rollin_mdp(env, rollin_type):
    """collect_data.py:83-111.  Returns states (H,2) int64, actions (H,5) f64, next_states (H,2)
    int64, rewards (H,) int64."""
    if rollin_type not in ("uniform", "expert"):
        raise NotImplementedError
    perm = None if getattr(env, "perm_index", None) is None else [env.perm_index]
    out = kernels.darkroom_rollin(np.asarray(env.goal)[None, :], env.dim, env.horizon, rollin_type, rng.next_key(), 0,
                                  perm, 0)
    return (out["context_states"][0].cpu().numpy().astype(np.int64),
            out["context_actions"][0].cpu().numpy().astype(np.float64),
            out["context_next_states"][0].cpu().numpy().astype(np.int64),
            out["context_rewards"][0, :, 0].cpu().numpy().astype(np.int64))


# --------------------------------------------------------------------------- generate_* -------
def _bandit_trajs(batch, n_samples):
    cs = batch["context_states"].cpu().numpy().astype(np.int64)
    ca = batch["context_actions"].cpu().numpy().astype(np.float64)
    cns = batch["context_next_states"].cpu().numpy().astype(np.int64)
    cr = batch["context_rewards"][:, :, 0].cpu().numpy().astype(np.float64)
    means = batch["means"].cpu().numpy().astype(np.float64)
    opt = batch["optimal_actions"].cpu().numpy().astype(np.float64)
    trajs = []
    for i in range(cs.shape[0]):
        for _ in range(n_samples):
            trajs.append({"query_state": np.array([1]), "optimal_action": opt[i], "context_states": cs[i],
                          "context_actions": ca[i], "context_next_states": cns[i], "context_rewards": cr[i],
                          "means": means[i]})
    return trajs


def generate_bandit_histories_from_envs(envs, n_hists, n_samples, cov, type):
    """collect_data.py:158-182, one launch per history index for all envs."""
    means = np.stack([np.asarray(e.means, dtype=np.float64) for e in envs])
    rtypes = {getattr(e, "type", "uniform") for e in envs}       # rewards follow env.type (env.transit, :45); the
    if len(rtypes) != 1 or len({float(e.var) for e in envs}) != 1:   # `type` argument is unused, like the reference's
        raise NotImplementedError("generate_bandit_histories_from_envs: all envs must share var and type")
    per_hist = []
    for _ in range(n_hists):
        b = collect_bandit(len(envs), means.shape[1], envs[0].H_context, envs[0].var, means=means, reward_type=next(iter(rtypes)))
        per_hist.append(_bandit_trajs(b, n_samples))
    trajs = []
    for i in range(len(envs)):           # env-major, then history, then sample -- the reference's order
        for j in range(n_hists):
            trajs.extend(per_hist[j][i * n_samples:(i + 1) * n_samples])
    for k, t in enumerate(trajs):        # keep the envs' own float64 means / opt_a objects, like the reference
        env = envs[k // (n_hists * n_samples)]
        t["means"], t["optimal_action"] = env.means, env.opt_a
    return trajs


def _bandit_trajs_host(host, means64, opt64, n_samples):
    """Traj dicts over views of the host arrays (no per-env copies)."""
    cs, ca, cns, cr = host["context_states"], host["context_actions"], host["context_next_states"], host["context_rewards"]
    q = np.array([1])
    trajs = []
    for i in range(cs.shape[0]):
        for _ in range(n_samples):
            trajs.append({"query_state": q, "optimal_action": opt64[i], "context_states": cs[i], "context_actions": ca[i],
                          "context_next_states": cns[i], "context_rewards": cr[i], "means": means64[i]})
    return trajs


def generate_bandit_histories(n_envs, dim, horizon, var, **kwargs):
    """collect_data.py:221-225.  Task draw and rollouts happen on the device; the histories come back in the reference's
    own dtypes (int64 states, float64 one-hot actions / rewards) through dpt_bandit_rollin_host_f64: 5 B per env-step
    over PCIe, expanded by the host cores straight into the returned arrays; the traj dicts hold views of them."""
    n_hists, n_samples = kwargs.get("n_hists", 1), kwargs.get("n_samples", 1)
    # the reference draws its envs with bandit_env.sample(dim, horizon, var) -- always type 'uniform' -- and never
    # reads kwargs['type'] (collect_data.py:158-182, 221-225); so do we
    key = rng.next_key()
    means, opt_idx, opt_a = kernels.bandit_sample_means(n_envs, dim, key, 0)
    means_host = means.cpu()
    means64, opt64 = means_host.numpy().astype(np.float64), opt_a.cpu().numpy().astype(np.float64)
    per_hist, scratch = [], None
    for j in range(n_hists):
        host = kernels.bandit_rollin_host_ref(means_host, horizon, float(var), rng.key_for(key, j + 1), 0)
        per_hist.append(_bandit_trajs_host(host, means64, opt64, n_samples))
    if n_hists == 1:
        return per_hist[0]
    trajs = []
    for i in range(n_envs):
        for j in range(n_hists):
            trajs.extend(per_hist[j][i * n_samples:(i + 1) * n_samples])
    return trajs


def _mdp_trajs(envs, out, n_samples):
    cs = out["context_states"].cpu().numpy().astype(np.int64)
    ca = out["context_actions"].cpu().numpy().astype(np.float64)
    cns = out["context_next_states"].cpu().numpy().astype(np.int64)
    cr = out["context_rewards"][:, :, 0].cpu().numpy().astype(np.int64)
    q = out["query_states"].cpu().numpy().astype(np.int64)
    oa = out["optimal_actions"].cpu().numpy().astype(np.float64)
    trajs = []
    for i, env in enumerate(envs):
        per_env = []
        for k in range(n_samples):
            t = {"query_state": q[i, k], "optimal_action": oa[i, k], "context_states": cs[i], "context_actions": ca[i],
                 "context_next_states": cns[i], "context_rewards": cr[i], "goal": env.goal}
            if hasattr(env, "perm_index"):
                t["perm_index"] = env.perm_index
            per_env.append(t)
        trajs.append(per_env)
    return trajs


def generate_mdp_histories_from_envs(envs, n_hists, n_samples, rollin_type):
    """collect_data.py:189-218 (one fused launch per history index)."""
    goals = np.stack([np.asarray(e.goal) for e in envs])
    perm = [e.perm_index for e in envs] if hasattr(envs[0], "perm_index") else None
    per_hist = [_mdp_trajs(envs, collect_darkroom(goals, envs[0].dim, envs[0].horizon, rollin_type, perm, n_samples),
                           n_samples) for _ in range(n_hists)]
    trajs = []
    for i in range(len(envs)):
        for j in range(n_hists):
            trajs.extend(per_hist[j][i])
    return trajs


def generate_darkroom_histories(goals, dim, horizon, **kwargs):
    """collect_data.py:290-293."""
    from .envs import darkroom_env
    envs = [darkroom_env.DarkroomEnv(dim, goal, horizon) for goal in goals]
    return generate_mdp_histories_from_envs(envs, **kwargs)


def generate_darkroom_permuted_histories(indices, dim, horizon, **kwargs):
    """collect_data.py:296-300."""
    from .envs import darkroom_env
    envs = [darkroom_env.DarkroomEnvPermuted(dim, index, horizon) for index in indices]
    return generate_mdp_histories_from_envs(envs, **kwargs)


# --------------------------------------------------------------------------- linear bandit ----
def rollin_linear_bandit_vec(envs):
    """collect_data.py:56-80: Thompson(prior_mean=0, prior_var=1, std=env.var, sample=True) driven
    through deploy_online_vec with include_meta=True -- here one fused launch."""
    from .ctrls.ctrl_bandit import ThompsonSamplingPolicy
    from .envs import bandit_env
    from .evals import eval_bandit
    H = envs[0].H_context
    thmp = ThompsonSamplingPolicy(envs[0], std=envs[0].var, sample=True, prior_mean=0.0, prior_var=1.0,
                                  warm_start=False, batch_size=len(envs))
    vec_env = bandit_env.BanditEnvVec(envs)
    _, meta = eval_bandit.deploy_online_vec(vec_env, thmp, H, include_meta=True)
    return (meta["context_states"], meta["context_actions"], meta["context_next_states"],
            meta["context_rewards"][:, :, 0])


def generate_linear_bandit_histories(n_envs, dim, lin_d, horizon, var, **kwargs):
    """collect_data.py:228-284.  (The reference reads n_hists / n_samples from module globals set in
    __main__, :241,:268 vs :363-364; here they are keyword arguments with the same names.)"""
    from .envs import bandit_env
    n_hists, n_samples = kwargs.get("n_hists", 1), kwargs.get("n_samples", 1)
    data_type = kwargs["data_type"]
    arms = np.random.RandomState(seed=1234).normal(size=(dim, lin_d)) / np.sqrt(lin_d)   # :230-231
    envs = [bandit_env.sample_linear(arms, horizon, var) for _ in range(n_envs)]
    if data_type == "thompson":
        per_hist = [rollin_linear_bandit_vec(envs) for _ in range(n_hists)]
    elif data_type == "uniform":
        means = np.stack([e.means for e in envs])
        per_hist = []
        for _ in range(n_hists):
            b = collect_bandit(n_envs, dim, horizon, var, means=means)
            per_hist.append((b["context_states"].cpu().numpy().astype(np.int64),
                             b["context_actions"].cpu().numpy().astype(np.float64),
                             b["context_next_states"].cpu().numpy().astype(np.int64),
                             b["context_rewards"][:, :, 0].cpu().numpy().astype(np.float64)))
    else:
        raise ValueError("Invalid data type")
    trajs = []
    for i, env in enumerate(envs):
        for j in range(n_hists):
            cs, ca, cns, cr = (x[i] for x in per_hist[j])
            for _ in range(n_samples):
                trajs.append({"query_state": np.array([1]), "optimal_action": env.opt_a, "context_states": cs,
                              "context_actions": ca, "context_next_states": cns, "context_rewards": cr,
                              "means": env.means, "arms": arms, "theta": env.theta, "var": env.var})
    return trajs


# --------------------------------------------------------------------------- dataset files -----
def build_datasets(env, n_envs, n_eval_envs, n_hists, n_samples, horizon, dim, var=0.0, cov=0.0, lin_d=2, out_dir="."):
    """The body of the reference's ``__main__`` (collect_data.py:352-485) as a function: train / test / eval
    trajectory lists for ``env`` in {'bandit', 'linear_bandit', 'darkroom_heldout'} (80 / 20 env split; darkroom:
    goals shuffled with RandomState(0), 80 / 20 goal split, each goal repeated ``n_envs // dim**2`` times, the eval set
    cycles the held-out goals to 100 envs), pickled under ``out_dir/datasets/`` with the reference's file names
    (utils.build_*_data_filename) and the reference's per-trajectory dict keys.  Returns the three paths.
    Seeding is the caller's (``dpt_b200.seed``), where the reference seeds ``np.random`` (:353-354)."""
    import os
    import pickle
    from . import utils
    n_train_envs = int(.8 * n_envs)
    n_test_envs = n_envs - n_train_envs
    config = {"n_hists": n_hists, "n_samples": n_samples, "horizon": horizon}
    if env == "bandit":
        config.update({"dim": dim, "var": var, "cov": cov, "type": "uniform"})
        sets = [generate_bandit_histories(n, **config) for n in (n_train_envs, n_test_envs, n_eval_envs)]
        name, n_eval_tag = utils.build_bandit_data_filename, n_eval_envs
    elif env == "linear_bandit":
        config.update({"dim": dim, "lin_d": lin_d, "var": var, "cov": cov, "data_type": "thompson"})
        sets = [generate_linear_bandit_histories(n, **config) for n in (n_train_envs, n_test_envs, n_eval_envs)]
        name, n_eval_tag = utils.build_linear_bandit_data_filename, n_eval_envs
    elif env == "darkroom_heldout":
        config.update({"dim": dim, "rollin_type": "uniform"})
        goals = np.array([[(j, i) for i in range(dim)] for j in range(dim)]).reshape(-1, 2)
        np.random.RandomState(seed=0).shuffle(goals)                                  # :408
        split = int(.8 * len(goals))
        train_goals, test_goals = goals[:split], goals[split:]
        eval_goals = np.array(test_goals.tolist() * int(100 // len(test_goals)))      # :413-414
        train_goals = np.repeat(train_goals, n_envs // (dim * dim), axis=0)
        test_goals = np.repeat(test_goals, n_envs // (dim * dim), axis=0)
        sets = [generate_darkroom_histories(g, **config) for g in (train_goals, test_goals, eval_goals)]
        name, n_eval_tag = utils.build_darkroom_data_filename, 100
    else:
        raise NotImplementedError
    paths = [os.path.join(out_dir, name(env, n, config, mode=m)) for m, n in ((0, n_envs), (1, n_envs), (2, n_eval_tag))]
    os.makedirs(os.path.join(out_dir, "datasets"), exist_ok=True)
    for path, trajs in zip(paths, sets):
        with open(path, "wb") as f:
            pickle.dump(trajs, f)
    return tuple(paths)
