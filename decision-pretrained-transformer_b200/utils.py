"""Host-side naming / conversion helpers of the reference's ``utils.py`` (the on-disk contract between
``collect_data.py``, ``train.py`` and ``eval.py``: dataset pickles and model checkpoints are located by these
file names).  Same function names, arguments and resulting strings; the miniworld builders are out of scope.
"""
import numpy as np
import torch

from . import kernels

_DATA = "datasets/trajs_{}.pkl"
_MODES = {0: "_train", 1: "_test", 2: "_eval"}
_MODEL_HEAD = (("_shuf", "shuffle"), ("_lr", "lr"), ("_do", "dropout"), ("_embd", "n_embd"), ("_layer", "n_layer"),
               ("_head", "n_head"), ("_envs", "n_envs"), ("_hists", "n_hists"), ("_samples", "n_samples"))


def _tagged(stem, config, fields):
    return stem + "".join(tag + str(config[key]) for tag, key in fields)


def _data_filename(env, n_envs, config, mode, fields, rollin_in_eval=False):
    name = env + "_envs" + str(n_envs)
    if mode != 2:                                             # eval sets carry no n_hists / n_samples
        name = _tagged(name, config, (("_hists", "n_hists"), ("_samples", "n_samples")))
    name = _tagged(name, config, fields)
    if mode == 2 and rollin_in_eval:
        name += "_" + config["rollin_type"]
    return _DATA.format(name + _MODES.get(mode, ""))


def build_bandit_data_filename(env, n_envs, config, mode):
    """utils.py:15-37.  mode 0 train / 1 test / 2 eval."""
    return _data_filename(env, n_envs, config, mode, (("_H", "horizon"), ("_d", "dim"), ("_var", "var"), ("_cov", "cov")))


def build_bandit_model_filename(env, config):
    """utils.py:41-60."""
    return _tagged(env, config, _MODEL_HEAD + (("_var", "var"), ("_cov", "cov"), ("_H", "horizon"), ("_d", "dim"), ("_seed", "seed")))


def build_linear_bandit_data_filename(env, n_envs, config, mode):
    """utils.py:62-85."""
    return _data_filename(env, n_envs, config, mode,
                          (("_H", "horizon"), ("_d", "dim"), ("_lind", "lin_d"), ("_var", "var"), ("_cov", "cov")))


def build_linear_bandit_model_filename(env, config):
    """utils.py:89-109."""
    return _tagged(env, config, _MODEL_HEAD + (("_var", "var"), ("_cov", "cov"), ("_H", "horizon"), ("_d", "dim"),
                                               ("_lind", "lin_d"), ("_seed", "seed")))


def build_darkroom_data_filename(env, n_envs, config, mode):
    """utils.py:111-133 (eval files also carry the rollin type)."""
    return _data_filename(env, n_envs, config, mode, (("_H", "horizon"), ("_d", "dim")), rollin_in_eval=True)


def build_darkroom_model_filename(env, config):
    """utils.py:136-153."""
    return _tagged(env, config, _MODEL_HEAD + (("_H", "horizon"), ("_d", "dim"), ("_seed", "seed")))


def convert_to_tensor(x, store_gpu=True):
    """utils.py:203-207: float32 tensor, on the CUDA device unless ``store_gpu`` is false."""
    t = torch.tensor(np.asarray(x)).float()
    return t.to(kernels._dev()) if store_gpu else t


def worker_init_fn(worker_id):
    """utils.py:6-10: per-DataLoader-worker torch / numpy seeds."""
    worker_seed = torch.initial_seed() % (2 ** 32) + worker_id
    torch.manual_seed(worker_seed)
    np.random.seed(int(worker_seed % (2 ** 32 - 1)))
