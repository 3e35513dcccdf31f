"""Bandit envs with the reference's interface (envs/bandit_env.py), backed by the CUDA path.

``BanditEnv`` / ``LinearBanditEnv`` keep the single-env API (reset / transit / step / deploy /
deploy_eval, attributes means, opt_a, opt_a_index, dim, dx, du, H_context, H, var, type);
``BanditEnvVec`` is where the work happens: it keeps the envs' means as one [N,d] device tensor
and steps all envs in one kernel launch (dpt_gpu_bandit_step) instead of a Python loop over env
objects (envs/bandit_env.py:98-105).  There is no CPU fallback.
"""
import numpy as np
import torch

from .. import kernels, rng
from .base_env import BaseEnv

_TYPES = {"uniform": 0, "bernoulli": 1}


def sample(dim, H, var, type="uniform"):
    """envs/bandit_env.py:10-18.  (The task draw itself is host-side numpy like the reference;
    use ``collect_data.generate_bandit_histories`` / ``kernels.bandit_sample_means`` for the
    batched device draw.)"""
    draw = {"uniform": lambda: np.random.uniform(0, 1, dim), "bernoulli": lambda: np.random.beta(1, 1, dim)}.get(type)
    if draw is None:
        raise NotImplementedError
    return BanditEnv(draw(), H, var=var, type=type)


def sample_linear(arms, H, var):
    """envs/bandit_env.py:21-25."""
    lin_d = arms.shape[1]
    theta = np.random.normal(0, 1, lin_d) / np.sqrt(lin_d)
    return LinearBanditEnv(theta, arms, H, var=var)


class BanditEnv(BaseEnv):
    """envs/bandit_env.py:28-82."""

    def __init__(self, means, H, var=0.0, type="uniform"):
        if type not in _TYPES:
            raise NotImplementedError
        self.means = np.asarray(means)
        self.dim = self.du = len(self.means)
        self.dx, self.state = 1, np.array([1])                     # the bandit "state" is the constant [1]
        self.opt_a_index = int(np.argmax(self.means))              # first maximum
        self.opt_a = np.eye(self.dim)[self.opt_a_index].reshape(self.means.shape)
        self.var, self.type, self.topk = var, type, False
        # the reference's naming (:44-47): H_context is the context length, an "episode" (H) is ONE pull
        self.H_context, self.H = H, 1
        self._vec = None                                           # lazily created 1-env device wrapper

    def get_arm_value(self, u):
        return np.sum(self.means * u)

    def reset(self):
        self.current_step = 0
        return self.state

    def _as_vec(self):
        if self._vec is None:
            self._vec = BanditEnvVec([self])
        return self._vec

    def transit(self, x, u):
        r = self._as_vec()._transit_device(torch.as_tensor(np.asarray(u)[None, :]))
        return self.state.copy(), float(r[0])

    def step(self, action):
        """:66-74"""
        if self.current_step >= self.H:
            raise ValueError("Episode has already ended")
        _, reward = self.transit(self.state, action)
        self.current_step += 1
        return self.state.copy(), reward, self.current_step >= self.H, {}

    def deploy_eval(self, ctrl):
        """:76-82 -- evaluation pulls are noise-free: var is zeroed for the duration of the rollout."""
        saved, self.var = self.var, 0.0
        try:
            return self.deploy(ctrl)
        finally:
            self.var = saved


class LinearBanditEnv(BanditEnv):
    """envs/bandit_env.py:158-197: means = arms @ theta."""

    def __init__(self, theta, arms, H, var=0.0):
        self.theta = theta
        self.arms = arms
        super().__init__(arms @ theta, H, var=var, type="uniform")


class BanditEnvVec(BaseEnv):
    """envs/bandit_env.py:85-153, one launch per step for all envs.

    ``envs``: list of BanditEnv-like objects (means, var, type, dx, du), as in the reference."""

    def __init__(self, envs, key=None):
        self._envs = envs
        self._num_envs = len(envs)
        self.dx = envs[0].dx
        self.du = envs[0].du
        self.device = kernels._dev()
        self.means = torch.as_tensor(np.stack([np.asarray(e.means, dtype=np.float64) for e in envs]),
                                     dtype=torch.float32).to(self.device)
        self._key = rng.next_key() if key is None else key
        self._draws = 0        # monotone step counter: part of the Philox counter
        # The reference's transit draws np.random.normal(0, var) per env even when var == 0 (envs/bandit_env.py:59).
        # Set True to burn the same N gaussians per step, so that controllers which draw from np.random
        # (Thompson sample=False, LinUCB first arm ...) see the reference's global-stream positions.
        self.np_random_compat = False
        self._done = [True] * self._num_envs

    @property
    def num_envs(self):
        return self._num_envs

    @property
    def envs(self):
        return self._envs

    def reset(self):
        for env in self._envs:
            env.current_step = 0
        return [env.state for env in self._envs]

    def _transit_device(self, actions, inject=None):
        env0 = self._envs[0]
        vars_ = {float(e.var) for e in self._envs}
        if len(vars_) != 1 or len({e.type for e in self._envs}) != 1:
            raise NotImplementedError("BanditEnvVec: all envs must share var and type")
        r = kernels.gpu_bandit_step(self.means, actions, float(env0.var), _TYPES[env0.type], self._key, 0,
                                    self._draws, inject=inject)
        self._draws += 1
        return r

    def step(self, actions):
        for env in self._envs:
            if env.current_step >= env.H:
                raise ValueError("Episode has already ended")
        r = self._transit_device(torch.as_tensor(np.asarray(actions))).cpu().numpy().astype(np.float64)
        if self.np_random_compat:
            np.random.standard_normal(self._num_envs)
        next_obs, dones = [], []
        for env in self._envs:
            env.current_step += 1
            next_obs.append(env.state.copy())
            dones.append(env.current_step >= env.H)
        return next_obs, list(r), dones, {}

    def deploy_eval(self, ctrl):
        tmp = [env.var for env in self._envs]
        for env in self._envs:
            env.var = 0.0
        res = self.deploy(ctrl)
        for env, var in zip(self._envs, tmp):
            env.var = var
        return res

    def deploy(self, ctrl):
        """envs/bandit_env.py:125-149: xs (N,1), us (N,du), xps (N,1), rs (N,)."""
        x = self.reset()
        xs, xps, us, rs = [], [], [], []
        done = False
        while not done:
            u = ctrl.act_numpy_vec(x)
            xs.append(x)
            us.append(u)
            x, r, done, _ = self.step(u)
            done = all(done)
            rs.append(r)
            xps.append(x)
        return np.concatenate(xs), np.concatenate(us), np.concatenate(xps), np.concatenate(rs)

    def get_arm_value(self, us):
        us = torch.as_tensor(np.asarray(us), dtype=torch.float32).to(self.device)
        return (self.means * us).sum(dim=1).cpu().numpy().astype(np.float64)
