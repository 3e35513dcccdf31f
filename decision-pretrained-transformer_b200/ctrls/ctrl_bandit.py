"""Bandit controllers with the reference's interface (ctrls/ctrl_bandit.py), backed by the CUDA path.

Every controller keeps the reference constructor and ``set_batch / set_batch_numpy_vec / set_env /
reset / act / act_numpy_vec``.  ``evals.eval_bandit.deploy_online_vec`` recognises these classes (via
``fused_spec``) and runs the whole H-step loop in one fused launch; the per-step methods below are
the compatible slow path (context -> dpt_arm_stats kernel -> decision on the device), for callers
that drive the loop themselves.  Statistics are float64 like the reference.
"""
import numpy as np
import torch

from .. import kernels, rng


class Controller:
    """ctrls/ctrl_bandit.py:11-19."""

    def set_batch(self, batch):
        self.batch = batch

    def set_batch_numpy_vec(self, batch):
        self.set_batch(batch)

    def set_env(self, env):
        self.env = env

    # -- helpers shared by the per-step paths ---------------------------------------------
    def _stats(self):
        """(sums f64 [N,d], counts f64 [N,d]) of the current batch via dpt_arm_stats."""
        acts, rews = self.batch["context_actions"], self.batch["context_rewards"]
        acts = torch.as_tensor(np.asarray(acts) if not torch.is_tensor(acts) else acts)
        rews = torch.as_tensor(np.asarray(rews) if not torch.is_tensor(rews) else rews)
        if acts.dim() == 2:
            acts, rews = acts[None], rews.reshape(1, -1, 1)
        N, h, d = acts.shape
        if h == 0:
            dev = kernels._dev()
            return torch.zeros((N, d), dtype=torch.float64, device=dev), torch.zeros((N, d), dtype=torch.float64, device=dev)
        sums, counts = kernels.arm_stats(acts.float(), rews.reshape(N, h, 1).float())
        return sums, counts.double()

    @staticmethod
    def _onehot(idx, d):
        a = np.zeros((idx.shape[0], d))
        a[np.arange(idx.shape[0]), idx] = 1.0
        return a


class OptPolicy(Controller):
    """ctrls/ctrl_bandit.py:22-38."""

    def __init__(self, env, batch_size=1):
        super().__init__()
        self.env = env
        self.batch_size = batch_size

    def reset(self):
        return

    def act(self, x):
        return self.env.opt_a

    def act_numpy_vec(self, x):
        return np.stack([env.opt_a for env in self.env], axis=0)

    def fused_spec(self):
        return dict(kind="opt")


class GreedyOptPolicy(Controller):
    """ctrls/ctrl_bandit.py:41-54: the context action with the largest observed reward."""

    def __init__(self, env):
        super().__init__()
        self.env = env

    def reset(self):
        return

    def act(self, x):
        rewards = torch.as_tensor(np.asarray(self.batch["context_rewards"].cpu() if torch.is_tensor(self.batch["context_rewards"])
                                             else self.batch["context_rewards"])).flatten()
        acts = self.batch["context_actions"]
        acts = acts.cpu().numpy() if torch.is_tensor(acts) else np.asarray(acts)
        self.a = acts[0][int(torch.argmax(rewards))]
        return self.a


class EmpMeanPolicy(Controller):
    """ctrls/ctrl_bandit.py:57-118."""

    def __init__(self, env, online=False, batch_size=1):
        super().__init__()
        self.env = env
        self.online = online
        self.batch_size = batch_size

    def reset(self):
        return

    def _decide(self):
        b, counts = self._stats()
        b_mean = b / torch.clamp(counts, min=1)
        i = torch.argmax(b_mean, dim=-1)
        if self.online:
            j = torch.argmin(counts, dim=-1)
            mask = counts.gather(1, j[:, None])[:, 0] == 0
            i = torch.where(mask, j, i)
        return i.cpu().numpy()

    def act(self, x):
        self.a = self._onehot(self._decide(), self.env.dim)[0]
        return self.a

    def act_numpy_vec(self, x):
        self.a = self._onehot(self._decide(), self.env.dim)
        return self.a

    def fused_spec(self):
        return dict(kind="emp", p0=1.0 if self.online else 0.0)


class UCBPolicy(Controller):
    """ctrls/ctrl_bandit.py:318-380.  The reference masks untried arms with a hard-coded
    ``np.arange(200)`` (:374, IndexError unless batch_size == 200); here the mask uses the
    batch size, which is the same thing at 200 and works everywhere else."""

    def __init__(self, env, const=1.0, batch_size=1):
        super().__init__()
        self.env = env
        self.const = const
        self.batch_size = batch_size

    def reset(self):
        return

    def _decide(self, override):
        b, counts = self._stats()
        b_mean = b / torch.clamp(counts, min=1)
        bounds = b_mean + self.const / torch.clamp(torch.sqrt(counts), min=1)
        i = torch.argmax(bounds, dim=-1)
        if override:
            j = torch.argmin(counts, dim=-1)
            mask = counts.gather(1, j[:, None])[:, 0] == 0
            i = torch.where(mask, j, i)
        return i.cpu().numpy()

    def act(self, x):          # the single-env path has no untried-arm override (:334-349)
        self.a = self._onehot(self._decide(False), self.env.dim)[0]
        return self.a

    def act_numpy_vec(self, x):
        self.a = self._onehot(self._decide(True), self.env.dim)
        return self.a

    def fused_spec(self):
        return dict(kind="ucb", p0=float(self.const))


class PessMeanPolicy(UCBPolicy):
    """ctrls/ctrl_bandit.py:255-314: lower confidence bound, no untried-arm override."""

    def _decide(self, override):
        b, counts = self._stats()
        bounds = b / torch.clamp(counts, min=1) - self.const / torch.clamp(torch.sqrt(counts), min=1)
        return torch.argmax(bounds, dim=-1).cpu().numpy()

    def fused_spec(self):
        return None


class ThompsonSamplingPolicy(Controller):
    """ctrls/ctrl_bandit.py:122-251."""

    def __init__(self, env, std=.1, sample=False, prior_mean=.5, prior_var=1 / 12.0, warm_start=False, batch_size=1):
        super().__init__()
        self.env = env
        self.std = std
        self.variance = std ** 2
        self.prior_mean = prior_mean
        self.prior_variance = prior_var
        self.batch_size = batch_size
        self.reset()
        self.sample = sample
        self.warm_start = warm_start

    def reset(self):
        shape = (self.batch_size, self.env.dim) if self.batch_size > 1 else (self.env.dim,)
        self.means = np.ones(shape) * self.prior_mean
        self.variances = np.ones(shape) * self.prior_variance
        self.counts = np.zeros(shape)

    def _posterior(self, batch):
        self.reset()
        self.batch = batch
        b, counts = self._stats()
        arm_means = torch.where(counts > 0, b / torch.clamp(counts, min=1), torch.zeros_like(b))
        pw = self.variance / (self.variance + counts * self.prior_variance)
        new_mean = pw * self.prior_mean + (1 - pw) * arm_means
        new_var = 1 / (1 / self.prior_variance + counts / self.variance)
        mask = counts > 0
        means = torch.where(mask, new_mean, torch.full_like(b, self.prior_mean))
        variances = torch.where(mask, new_var, torch.full_like(b, self.prior_variance))
        self.means = means.cpu().numpy().reshape(self.means.shape)
        self.variances = variances.cpu().numpy().reshape(self.means.shape)
        self.counts = counts.cpu().numpy().reshape(self.means.shape)

    def set_batch(self, batch):
        self._posterior(batch)

    def set_batch_numpy_vec(self, batch):
        self._posterior(batch)

    def update_posterior(self, c, arm_rewards):
        """:184-194 (single env): conjugate update of arm ``c`` from its observed rewards."""
        n = self.counts[c]
        if n > 0:
            arm_mean = np.mean(arm_rewards)
            prior_weight = self.variance / (self.variance + (n * self.prior_variance))
            self.means[c] = prior_weight * self.prior_mean + (1 - prior_weight) * arm_mean
            self.variances[c] = 1 / (1 / self.prior_variance + n / self.variance)

    def update_posterior_all(self, arm_means):
        """:196-203 (batched): the same update for every (env, arm) with a positive count."""
        prior_weight = self.variance / (self.variance + (self.counts * self.prior_variance))
        new_mean = prior_weight * self.prior_mean + (1 - prior_weight) * arm_means
        new_variance = 1 / (1 / self.prior_variance + self.counts / self.variance)
        mask = (self.counts > 0)
        self.means[mask] = new_mean[mask]
        self.variances[mask] = new_variance[mask]

    def _draw(self):
        if self.sample:
            values = np.random.normal(self.means, np.sqrt(self.variances))
            return np.argmax(values, axis=-1)
        values = np.stack([np.random.normal(self.means, np.sqrt(self.variances)) for _ in range(100)], axis=-2)
        amax = np.argmax(values, axis=-1)
        if amax.ndim == 1:
            return np.argmax(np.bincount(amax, minlength=self.env.dim))
        return np.argmax(np.array([np.bincount(am, minlength=self.env.dim) for am in amax]), axis=-1)

    def act(self, x):
        i = int(self._draw())
        if self.sample and self.warm_start:
            j = int(np.argmin(self.counts))
            if self.counts[j] == 0:
                i = j
        a = np.zeros(self.env.dim)
        a[i] = 1.0
        self.a = a
        return self.a

    def act_numpy_vec(self, x):
        self.a = self._onehot(np.atleast_1d(self._draw()), self.env.dim)
        return self.a

    def fused_spec(self):
        if not self.sample:
            return None
        return dict(kind="thompson", p0=float(self.std), p1=float(self.prior_mean), p2=float(self.prior_variance))


class LinUCBPolicy(OptPolicy):
    """ctrls/ctrl_bandit.py:447-528."""

    def __init__(self, env, const=1.0, batch_size=1):
        super().__init__(env)
        self.rand = True
        self.const = const
        self.arms = env.arms
        self.d = self.arms.shape[1]
        self.dim = env.dim
        self.theta = np.zeros(self.d)
        self.init_cov = 1.0 * np.eye(self.d)
        self.batch_size = batch_size

    def act_numpy_vec(self, x):
        actions_batch, rewards_batch = self.batch["context_actions"], self.batch["context_rewards"]
        if len(rewards_batch[0]) < 1:
            idx = np.random.choice(np.arange(self.dim), size=self.batch_size)
            return self._onehot(idx, self.dim)
        dev = kernels._dev()
        acts = torch.as_tensor(np.asarray(actions_batch)).to(dev).double()
        rews = torch.as_tensor(np.asarray(rewards_batch)).to(dev).double().reshape(acts.shape[0], -1, 1)
        arms = torch.as_tensor(self.arms).to(dev).double()
        A = arms[acts.argmax(-1)]                                  # [N,h,lin_d]
        cov = torch.eye(self.d, device=dev, dtype=torch.float64) + A.transpose(1, 2) @ A
        cov_inv = torch.linalg.inv(cov)
        theta = (cov_inv @ A.transpose(1, 2) @ rews)[:, :, 0]      # [N,lin_d]
        q = torch.einsum("ai,nij,aj->na", arms, cov_inv, arms)
        vals = theta @ arms.T + self.const * torch.sqrt(q)
        return self._onehot(vals.argmax(-1).cpu().numpy(), self.dim)

    def fused_spec(self):
        return dict(kind="linucb", p0=float(self.const), arms=np.asarray(self.arms, dtype=np.float64))


class BanditTransformerController(Controller):
    """ctrls/ctrl_bandit.py:383-444.  ``model`` is this package's ``models.net.Transformer``.

    Per-step path: the context is handed to the fused dense forward (no per-step re-upload of
    float64 arrays as in :396-401 when it already lives on the device).  Fused path:
    ``deploy_online_vec`` runs the KV-cached decode + sampling + env step loop in one launch."""

    def __init__(self, model, sample=False, batch_size=1):
        self.model = model
        self.du = model.config["action_dim"]
        self.dx = model.config["state_dim"]
        self.H = model.horizon
        self.sample = sample
        self.batch_size = batch_size
        self.zeros = torch.zeros(batch_size, self.dx ** 2 + self.du + 1, device=kernels._dev())

    def set_env(self, env):
        return

    def set_batch_numpy_vec(self, batch):
        dev = kernels._dev()
        self.set_batch({k: torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).float().to(dev)
                        for k, v in batch.items()})

    def _logits(self, x):
        self.batch["zeros"] = self.zeros
        states = torch.as_tensor(np.array(x)).float().to(self.zeros.device)
        if states.dim() == 1:
            states = states[None, :]
        self.batch["query_states"] = states
        return self.model(self.batch).detach().cpu().numpy().astype(np.float64)

    def _pick(self, a):
        if self.sample:
            e = np.exp(a - a.max(axis=-1, keepdims=True))
            probs = e / e.sum(axis=-1, keepdims=True)
            return np.array([np.random.choice(np.arange(self.du), p=p) for p in probs])
        return np.argmax(a, axis=-1)

    def act(self, x):
        return self._onehot(self._pick(self._logits(x)), self.du)[0]

    def act_numpy_vec(self, x):
        return self._onehot(self._pick(self._logits(x)), self.du)

    def fused_spec(self):
        if self.dx != 1 or not hasattr(self.model, "online_loop"):
            return None
        return dict(kind="transformer")

    def fused_online_loop(self, means, horizon, var, key, env_id0, include_meta, regret, inject, dump, reward_type="uniform"):
        return self.model.online_loop(means, horizon, var, self.sample, key, env_id0, include_meta, regret, inject, dump,
                                      reward_type)
