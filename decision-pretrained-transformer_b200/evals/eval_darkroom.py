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


def deploy_online_vec(vec_env, controller, Heps, H, horizon):
    """evals/eval_darkroom.py:20-84.  Returns per-episode returns [num_envs, Heps]."""
    assert H % horizon == 0
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
