"""Darkroom online evaluation with the reference's interface (evals/eval_darkroom.py:20-121).

Episode-level loop: the first H // horizon episodes grow the context, later ones slide it by one
episode.  Within an episode the context is fixed and only the query state (position 0 of the
sequence) changes, so every step is a dense forward over the whole context (no K/V reuse across
steps).  Context tensors stay on the device; env steps are one dpt_darkroom_step
launch, forwards are dpt_gpt2_forward.
"""
import numpy as np
import torch

from .. import kernels
from ..ctrls.ctrl_darkroom import DarkroomOptPolicy, DarkroomTransformerController  # noqa: F401
from ..envs.darkroom_env import DarkroomEnv, DarkroomEnvPermuted, DarkroomEnvVec


def deploy_online_vec_device(vec_env, model, Heps, H, horizon, sample=True, seed=None, env_id0=0, inject_u=None, dump=False):
    """The same loop with everything on the device and TWO launches per episode.

    Within an episode the context is fixed, so the controller's logits depend only on the query state.
    Instead of `horizon` sequential forwards per episode, ONE batched dense forward evaluates all
    dim*dim query states of every env against its context (ctx_share = dim*dim: the states of an env
    share one context row), and a rollout kernel then plays the episode from that logits table
    (softmax + categorical draw + grid transition per step).  Results are identical to the step-by-step
    loop given the same uniforms (``inject_u`` f64 [Heps*horizon, N]).  Returns a dict with
    ``returns`` [N,Heps] (device) and the final context tensors [, ``u``]."""
    from .. import rng
    assert H % horizon == 0
    ctx_rollouts = H // horizon
    dev = kernels._dev()
    n = vec_env.num_envs
    dim = vec_env.envs[0].dim
    goals = torch.as_tensor(np.stack([e.goal for e in vec_env.envs]), dtype=torch.int32).to(dev)
    perms = [getattr(e, "perm_index", None) for e in vec_env.envs]
    perm = None if perms[0] is None else torch.as_tensor(perms, dtype=torch.int32).to(dev)
    key = rng.next_key() if seed is None else seed
    g = torch.arange(dim, device=dev, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)          # index x*dim + y
    q = grid.repeat(n, 1).contiguous()
    names = ("states", "actions", "next_states", "rewards")
    ctx = {k: torch.zeros((n, 0, w), device=dev) for k, w in zip(names, (2, 5, 2, 1))}
    inj = None if inject_u is None else torch.as_tensor(np.asarray(inject_u), dtype=torch.float64).to(dev)
    rets, us = [], []
    was_test = model.test
    model.test = True
    for ep in range(Heps):
        logits = model.forward({"query_states": q, "context_states": ctx["states"], "context_actions": ctx["actions"],
                                "context_next_states": ctx["next_states"], "context_rewards": ctx["rewards"]},
                               ctx_share=dim * dim)
        ro = kernels.darkroom_policy_rollout(logits.view(n, dim * dim, 5), goals, dim, horizon, sample, key, env_id0, ep, perm,
                                             None if inj is None else inj[ep * horizon:(ep + 1) * horizon].contiguous(), dump)
        rets.append(ro["returns"])
        if dump:
            us.append(ro["u"])
        drop = horizon if ep >= ctx_rollouts else 0          # slide the window by one episode (:73-82)
        ctx = {k: torch.cat((ctx[k][:, drop:], ro[k]), dim=1).contiguous() for k in names}
    model.test = was_test
    out = {"returns": torch.stack(rets, dim=1), "context": ctx}
    if dump:
        out["u"] = torch.cat(us, dim=0)
    return out


def deploy_online_vec(vec_env, controller, Heps, H, horizon):
    """evals/eval_darkroom.py:20-84.  Returns per-episode returns [num_envs, Heps]."""
    assert H % horizon == 0
    if (isinstance(vec_env, DarkroomEnvVec) and isinstance(controller, DarkroomTransformerController)
            and getattr(controller, "fused", True) and controller.temp == 1.0):
        out = deploy_online_vec_device(vec_env, controller.model, Heps, H, horizon, sample=controller.sample)
        return out["returns"].cpu().numpy().astype(np.float64)
    ctx_rollouts = H // horizon
    dev = kernels._dev()
    n = vec_env.num_envs
    f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).to(dev)   # noqa: E731
    cs = torch.zeros((n, ctx_rollouts, horizon, vec_env.state_dim), device=dev)
    ca = torch.zeros((n, ctx_rollouts, horizon, vec_env.action_dim), device=dev)
    cns = torch.zeros((n, ctx_rollouts, horizon, vec_env.state_dim), device=dev)
    cr = torch.zeros((n, ctx_rollouts, horizon, 1), device=dev)
    cum_means = []
    for i in range(Heps):
        k = min(i, ctx_rollouts)
        controller.set_batch({"context_states": cs[:, :k].reshape(n, -1, vec_env.state_dim),
                              "context_actions": ca[:, :k].reshape(n, -1, vec_env.action_dim),
                              "context_next_states": cns[:, :k].reshape(n, -1, vec_env.state_dim),
                              "context_rewards": cr[:, :k].reshape(n, -1, 1)})
        s, a, ns, r = vec_env.deploy_eval(controller)
        cum_means.append(np.sum(r, axis=-1))
        if i < ctx_rollouts:
            cs[:, i], ca[:, i], cns[:, i], cr[:, i] = f(s), f(a), f(ns), f(r[:, :, None])
        else:   # slide the window by one episode (:73-82)
            cs = torch.cat((cs[:, 1:], f(s)[:, None]), dim=1)
            ca = torch.cat((ca[:, 1:], f(a)[:, None]), dim=1)
            cns = torch.cat((cns[:, 1:], f(ns)[:, None]), dim=1)
            cr = torch.cat((cr[:, 1:], f(r[:, :, None])[:, None]), dim=1)
    return np.stack(cum_means, axis=1)


def online(eval_trajs, model, Heps, H, n_eval, dim, horizon, permuted=False):
    """evals/eval_darkroom.py:87-121 without plotting: returns (returns [n_eval,Heps], mean, sem)."""
    assert H % horizon == 0
    envs = [DarkroomEnvPermuted(dim, eval_trajs[i]["perm_index"], horizon) if permuted
            else DarkroomEnv(dim, eval_trajs[i]["goal"], horizon) for i in range(n_eval)]
    ctrl = DarkroomTransformerController(model, batch_size=n_eval, sample=True)
    all_means = deploy_online_vec(DarkroomEnvVec(envs), ctrl, Heps, H, horizon)
    return all_means, all_means.mean(0), all_means.std(0, ddof=1) / np.sqrt(n_eval)


def offline_device(eval_trajs, model, n_eval, H, dim, permuted=False, seed=None, inject_u=None, dump=False):
    """The learner half of ``offline`` with everything on the device: the context of every env is its eval
    trajectory (fixed for the whole episode), so ONE batched dense forward gives the logits of all dim*dim
    query states of every env and ONE rollout launch per policy plays the H-step episode from that table
    (sampled with the Philox stream or ``inject_u`` f64 [H, N]; greedy = argmax).  Returns a dict of device
    tensors: ``returns_sample`` / ``returns_greedy`` [n_eval] and the two rollouts."""
    from .. import rng
    dev = kernels._dev()
    trajs = [eval_trajs[i] for i in range(n_eval)]
    f = lambda k, tail: torch.as_tensor(np.stack([np.asarray(t[k], dtype=np.float64).reshape((-1,) + tail) for t in trajs]),  # noqa: E731
                                        dtype=torch.float32).to(dev)
    ctx = {"context_states": f("context_states", (2,)), "context_actions": f("context_actions", (5,)),
           "context_next_states": f("context_next_states", (2,)), "context_rewards": f("context_rewards", (1,))}
    if permuted:
        goals = torch.full((n_eval, 2), dim - 1, dtype=torch.int32, device=dev)            # envs/darkroom_env.py:92
        perm = torch.as_tensor([int(t["perm_index"]) for t in trajs], dtype=torch.int32).to(dev)
    else:
        goals = torch.as_tensor(np.stack([np.asarray(t["goal"]) for t in trajs]), dtype=torch.int32).to(dev)
        perm = None
    g = torch.arange(dim, device=dev, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)              # index x*dim + y
    was_test = model.test
    model.test = True
    logits = model.forward(dict(ctx, query_states=grid.repeat(n_eval, 1).contiguous()), ctx_share=dim * dim)
    model.test = was_test
    table = logits.view(n_eval, dim * dim, 5)
    key = rng.next_key() if seed is None else seed
    inj = None if inject_u is None else torch.as_tensor(np.asarray(inject_u), dtype=torch.float64).to(dev)
    smp = kernels.darkroom_policy_rollout(table, goals, dim, H, True, key, 0, 0, perm, inj, dump)
    grd = kernels.darkroom_policy_rollout(table, goals, dim, H, False, key, 0, 1, perm)
    return {"returns_sample": smp["returns"], "returns_greedy": grd["returns"], "sample": smp, "greedy": grd, "logits": table}


def offline(eval_trajs, model, n_eval, H, dim, permuted=False):
    """evals/eval_darkroom.py:124-190 without the bar plot.  Returns {'Opt', 'Learner', 'Learner (greedy)'}:
    per-env returns of an H-step episode for the optimal policy and for the transformer (sampled / greedy)
    conditioned on each eval trajectory's own context."""
    trajs = [eval_trajs[i] for i in range(n_eval)]
    if permuted:
        goals = np.full((n_eval, 2), dim - 1)
        perm = np.array([int(t["perm_index"]) for t in trajs])
    else:
        goals, perm = np.stack([np.asarray(t["goal"]) for t in trajs]), None
    # DarkroomOptPolicy from the reset state (:151-155) is the 'expert' rollin of the rollout kernel
    opt = kernels.darkroom_rollin(goals, dim, H, "expert", 0, 0, perm, 1)
    rs_opt = opt["context_rewards"][:, :, 0].sum(dim=1)
    out = offline_device(eval_trajs, model, n_eval, H, dim, permuted)
    return {"Opt": rs_opt.cpu().numpy().astype(np.float64), "Learner": out["returns_sample"].cpu().numpy().astype(np.float64),
            "Learner (greedy)": out["returns_greedy"].cpu().numpy().astype(np.float64)}
