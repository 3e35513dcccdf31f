"""The reference's abstract env interface (envs/base_env.py:8-48) without the gym dependency -- gym is only a
base class there.  Subclasses provide ``reset`` / ``transit`` / ``step``; ``deploy`` is the generic single-env
rollout that the vectorised envs override with batched device code."""
import numpy as np


def rollout(env, policy, all_done=bool):
    """Play ``policy`` on ``env`` from a reset until the env reports done (``all_done`` reduces a vectorised env's
    per-env flags to one).

    Yields one ``(observation, action, next_observation, reward)`` tuple per step, in the order the reference's
    loop appends them (envs/base_env.py:31-41): the action is chosen from the observation BEFORE the step."""
    observation, finished = env.reset(), False
    while not finished:
        action = policy.act(observation)
        successor, reward, flags, _info = env.step(action)
        finished = all_done(flags)
        yield observation, action, successor, reward
        observation = successor


class BaseEnv:
    """Interface only; every method a subclass does not provide raises like an abstract method would."""

    def _abstract(self, *_args, **_kwargs):
        raise NotImplementedError

    reset = transit = step = _abstract

    def render(self, mode="human"):
        """The reference's envs never render (envs/base_env.py:18-19)."""

    def deploy(self, ctrl):
        """envs/base_env.py:24-48: four stacked arrays (obs, acts, next_obs, rews), one row per step."""
        columns = tuple(zip(*rollout(self, ctrl))) or ((), (), (), ())
        return tuple(np.array(col) for col in columns)

    def deploy_eval(self, ctrl):
        """envs/base_env.py:21-22: evaluation rollouts are plain rollouts unless a subclass removes noise."""
        return self.deploy(ctrl)
