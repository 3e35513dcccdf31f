"""Darkroom controllers with the reference's interface (ctrls/ctrl_darkroom.py)."""
import numpy as np
import torch

from .. import kernels
from .ctrl_bandit import Controller


class DarkroomOptPolicy(Controller):
    """ctrls/ctrl_darkroom.py:10-20."""

    def __init__(self, env):
        super().__init__()
        self.env = env
        self.goal = env.goal

    def reset(self):
        return

    def act(self, state):
        return self.env.opt_action(state)


class DarkroomTransformerController(Controller):
    """ctrls/ctrl_darkroom.py:23-66: logits = model(batch) with the current states as query
    (dpt_gpt2_forward), then softmax(logits / temp) + np.random.choice per env, or argmax."""

    def __init__(self, model, batch_size=1, sample=False):
        self.model = model
        self.state_dim = model.config["state_dim"]
        self.action_dim = model.config["action_dim"]
        self.horizon = model.horizon
        self.zeros = torch.zeros(batch_size, self.state_dim ** 2 + self.action_dim + 1, device=kernels._dev())
        self.sample = sample
        self.temp = 1.0
        self.batch_size = batch_size

    def act(self, state):
        self.batch["zeros"] = self.zeros
        states = torch.as_tensor(np.array(state)).float().to(self.zeros.device)
        if self.batch_size == 1:
            states = states[None, :]
        self.batch["query_states"] = states
        actions = self.model(self.batch).cpu().numpy().astype(np.float64)
        if self.sample:
            z = actions / self.temp
            e = np.exp(z - z.max(axis=-1, keepdims=True))
            probs = e / e.sum(axis=-1, keepdims=True)
            action_indices = [np.random.choice(np.arange(self.action_dim), p=p) for p in probs]
        else:
            action_indices = np.argmax(actions, axis=-1)
        out = np.zeros((self.batch_size, self.action_dim))
        out[np.arange(self.batch_size), action_indices] = 1.0
        return out[0] if self.batch_size == 1 else out
