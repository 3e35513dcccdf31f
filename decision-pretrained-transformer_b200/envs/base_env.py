"""BaseEnv: the reference's abstract env interface (envs/base_env.py:8-48), without the gym
dependency (gym is only a base class there).  ``deploy`` is the generic single-env loop."""
import numpy as np


class BaseEnv:
    def reset(self):
        raise NotImplementedError

    def transit(self, state, action):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self, mode="human"):
        pass

    def deploy_eval(self, ctrl):
        return self.deploy(ctrl)

    def deploy(self, ctrl):
        """envs/base_env.py:24-48: roll ``ctrl`` until done; returns (obs, acts, next_obs, rews)."""
        ob = self.reset()
        obs, acts, next_obs, rews = [], [], [], []
        done = False
        while not done:
            act = ctrl.act(ob)
            obs.append(ob)
            acts.append(act)
            ob, rew, done, _ = self.step(act)
            rews.append(rew)
            next_obs.append(ob)
        return np.array(obs), np.array(acts), np.array(next_obs), np.array(rews)
