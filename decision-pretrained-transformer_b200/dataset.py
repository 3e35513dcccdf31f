"""Dataset with the reference's interface (dataset.py:11-91): the hand-off between collection and training.

``Dataset(path, config)`` reads the pickled list of trajectory dicts exactly like the reference;
``Dataset.from_trajs`` / ``Dataset.from_batch`` build the same object straight from a trajectory list or
from the DEVICE-RESIDENT batch returned by ``collect_data.collect_bandit`` / ``collect_darkroom`` -- the
latter skips the host round trip and the (16 GB at BASELINE config 5) pickle altogether.  Items have the
reference's keys, fp32 dtype and shapes (rewards [H,1], zeros [dx^2 + du + 1]), with the same optional
per-item context shuffle.
"""
import pickle

import numpy as np
import torch


def convert_to_tensor(x, store_gpu=True):
    """utils.py:193-197."""
    t = torch.as_tensor(np.asarray(x)).float()
    return t.to("cuda") if store_gpu and torch.cuda.is_available() else t


class Dataset(torch.utils.data.Dataset):
    def __init__(self, path, config):
        if not isinstance(path, list):
            path = [path]
        trajs = []
        for p in path:
            with open(p, "rb") as f:
                trajs += pickle.load(f)
        self._init_from_trajs(trajs, config)

    @classmethod
    def from_trajs(cls, trajs, config):
        self = cls.__new__(cls)
        self._init_from_trajs(trajs, config)
        return self

    @classmethod
    def from_batch(cls, batch, config):
        """``batch``: dict of device tensors with context_* [N,H,.], query_states [N,dx] (optional for
        bandits: the constant [1]) and optimal_actions [N,du] (or [N,S,du]: the first sample is used)."""
        self = cls.__new__(cls)
        self._set_config(config)
        self.trajs = None
        n = batch["context_actions"].shape[0]
        dev = batch["context_actions"].device
        q = batch.get("query_states")
        if q is None:
            q = torch.ones((n, config["state_dim"]), device=dev)
        oa = batch["optimal_actions"]
        self.dataset = {
            "query_states": (q[:, 0] if q.dim() == 3 else q).float(),
            "optimal_actions": (oa[:, 0] if oa.dim() == 3 else oa).float(),
            "context_states": batch["context_states"].float(),
            "context_actions": batch["context_actions"].float(),
            "context_next_states": batch["context_next_states"].float(),
            "context_rewards": batch["context_rewards"].float().reshape(n, -1, 1),
        }
        self.zeros = torch.zeros(config["state_dim"] ** 2 + config["action_dim"] + 1, device=dev)
        return self

    def _set_config(self, config):
        self.shuffle = config["shuffle"]
        self.horizon = config["horizon"]
        self.store_gpu = config["store_gpu"]
        self.config = config

    def _init_from_trajs(self, trajs, config):
        self._set_config(config)
        self.trajs = trajs
        g = lambda k: np.array([t[k] for t in trajs])   # noqa: E731
        context_rewards = g("context_rewards")
        if len(context_rewards.shape) < 3:
            context_rewards = context_rewards[:, :, None]
        self.dataset = {
            "query_states": convert_to_tensor(g("query_state"), store_gpu=self.store_gpu),
            "optimal_actions": convert_to_tensor(g("optimal_action"), store_gpu=self.store_gpu),
            "context_states": convert_to_tensor(g("context_states"), store_gpu=self.store_gpu),
            "context_actions": convert_to_tensor(g("context_actions"), store_gpu=self.store_gpu),
            "context_next_states": convert_to_tensor(g("context_next_states"), store_gpu=self.store_gpu),
            "context_rewards": convert_to_tensor(context_rewards, store_gpu=self.store_gpu),
        }
        self.zeros = convert_to_tensor(np.zeros(config["state_dim"] ** 2 + config["action_dim"] + 1), store_gpu=self.store_gpu)

    def __len__(self):
        return len(self.dataset["query_states"])

    def __getitem__(self, index):
        res = {k: self.dataset[k][index] for k in ("context_states", "context_actions", "context_next_states", "context_rewards",
                                                   "query_states", "optimal_actions")}
        res["zeros"] = self.zeros
        if self.shuffle:
            perm = torch.randperm(self.horizon)
            for k in ("context_states", "context_actions", "context_next_states", "context_rewards"):
                res[k] = res[k][perm]
        return res
