"""Darkroom envs with the reference's interface (envs/darkroom_env.py), backed by the CUDA path.

Single-env objects keep the reference API (sample_state, sample_action, reset, transit, step,
opt_action, get_obs); ``DarkroomEnvVec`` steps all envs in one launch (dpt_darkroom_step) instead
of a Python loop (envs/darkroom_env.py:126-133).  Transitions are integer arithmetic on the
device -- bit-exact against the reference.
"""
import itertools

import numpy as np
import torch

from .. import kernels
from .base_env import BaseEnv, rollout


class _Discrete:
    def __init__(self, n):
        self.n = n


class DarkroomEnv(BaseEnv):
    """envs/darkroom_env.py:12-82."""

    def __init__(self, dim, goal, horizon):
        self.dim = dim
        self.goal = np.array(goal)
        self.horizon = horizon
        self.state_dim = 2
        self.action_dim = 5
        self.action_space = _Discrete(self.action_dim)
        self._perm_index = None

    def sample_state(self):
        """:23-24 -- a uniform grid cell from the host ``np.random`` stream (the fused collection draws on the device)."""
        return np.random.randint(0, self.dim, 2)

    def sample_action(self):
        """:26-30 -- a uniform one-hot action."""
        return np.eye(self.action_space.n)[np.random.randint(0, 5)]

    def reset(self):
        """:32-35 -- every episode starts in the corner (0, 0)."""
        self.current_step, self.state = 0, np.array([0, 0])
        return self.state

    def _perm_arg(self):
        return None if self._perm_index is None else [self._perm_index]

    def transit(self, state, action):
        ns, r = kernels.darkroom_step(np.asarray(state)[None, :], np.asarray(action, dtype=np.float32)[None, :],
                                      self.goal[None, :], self.dim, self._perm_arg())
        return ns[0].cpu().numpy().astype(np.int64), int(r[0])

    def step(self, action):
        """:57-64"""
        if self.current_step >= self.horizon:
            raise ValueError("Episode has already ended")
        self.state, reward = self.transit(self.state, action)
        self.current_step += 1
        return self.state.copy(), reward, self.current_step >= self.horizon, {}

    def get_obs(self):
        return self.state.copy()

    def opt_action(self, state):
        a = kernels.darkroom_opt_action(np.asarray(state)[None, :], self.goal[None, :], self._perm_arg())
        return a[0].cpu().numpy().astype(np.float64)


class DarkroomEnvPermuted(DarkroomEnv):
    """envs/darkroom_env.py:85-111: goal in the far corner, actions relabelled by a permutation."""

    def __init__(self, dim, perm_index, H):
        super().__init__(dim, np.array([dim - 1, dim - 1]), H)
        assert perm_index < 120     # 5! permutations in darkroom
        self.perm_index = perm_index
        self._perm_index = perm_index
        self.perm = list(itertools.permutations(np.arange(self.action_space.n)))[perm_index]


class DarkroomEnvVec(BaseEnv):
    """envs/darkroom_env.py:114-175 with one launch per step."""

    def __init__(self, envs):
        self._envs = envs
        self._num_envs = len(envs)
        self.device = kernels._dev()
        self._goals = torch.as_tensor(np.stack([e.goal for e in envs]), dtype=torch.int32).to(self.device)
        perms = [getattr(e, "perm_index", None) for e in envs]
        self._perm = None if perms[0] is None else torch.as_tensor(perms, dtype=torch.int32).to(self.device)
        self._dim = envs[0].dim

    @property
    def num_envs(self):
        return self._num_envs

    @property
    def envs(self):
        return self._envs

    @property
    def state_dim(self):
        return self._envs[0].state_dim

    @property
    def action_dim(self):
        return self._envs[0].action_dim

    def reset(self):
        return [env.reset() for env in self._envs]

    def step(self, actions):
        for env in self._envs:
            if env.current_step >= env.horizon:
                raise ValueError("Episode has already ended")
        states = np.stack([env.state for env in self._envs])
        ns, r = kernels.darkroom_step(states, np.asarray(actions, dtype=np.float32), self._goals, self._dim, self._perm)
        ns, r = ns.cpu().numpy().astype(np.int64), r.cpu().numpy().astype(np.int64)
        next_obs, dones = [], []
        for i, env in enumerate(self._envs):
            env.state = ns[i]
            env.current_step += 1
            next_obs.append(env.state.copy())
            dones.append(env.current_step >= env.horizon)
        return next_obs, [int(x) for x in r], dones, {}

    def deploy(self, ctrl):
        """envs/darkroom_env.py:151-175: ``ctrl.act`` sees the list of all envs' states; the four results are
        stacked env-major: obs [N,horizon,2], acts [N,horizon,5], next_obs [N,horizon,2], rews [N,horizon]."""
        columns = tuple(zip(*rollout(self, ctrl, all_done=all)))
        return tuple(np.stack(col, axis=1) for col in columns)
