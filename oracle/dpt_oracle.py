"""numpy restatement of the reference's rollout hot path -- TEST INFRASTRUCTURE.

Each function cites the reference file:line it follows (paths relative to the
reference root).  All arithmetic is float64 / int64 like the reference.  Random
draws go through a *noise source* so that the same code can

  * draw from numpy's legacy global stream in exactly the reference's order
    (``GlobalNoise`` -- used by ``oracle/make_golden.py`` to pin this file
    bit-for-bit against the live reference and to record the consumed noise), or
  * replay recorded / device-generated noise (``ReplayNoise`` -- used by the
    parity tests: the CUDA path and this oracle consume identical noise).

Parity pin: pinned against the live reference by ``oracle/make_golden.py``
(outputs identical under the same ``np.random.seed``); golden vectors in
``tests/golden/*.npz``.
"""
import itertools

import numpy as np

COV_GRID = [0.0, .1, .2, .3, .4, .5, .6, .7, .8, .9, 1.0]  # collect_data.py:30
DARKROOM_PERMS = list(itertools.permutations(range(5)))     # envs/darkroom_env.py:96-98


# ----------------------------------------------------------------------------
# noise sources
# ----------------------------------------------------------------------------
class GlobalNoise:
    """numpy legacy global stream, same calls as the reference; records every draw."""

    def __init__(self, record=True):
        self.rec = {}
        self.record = record

    def _r(self, name, v):
        if self.record:
            self.rec.setdefault(name, []).append(np.array(v))
        return v

    def choice_index(self, n, name):            # np.random.choice(a) without p -> randint(0, len(a))
        return self._r(name, int(np.random.choice(np.arange(n))))

    def choice_index_vec(self, n, size, name):  # ctrls/ctrl_bandit.py:497
        return self._r(name, np.random.choice(np.arange(n), size=size))

    def dirichlet(self, d, name):               # collect_data.py:32
        return self._r(name, np.random.dirichlet(np.ones(d)))

    def uniform01(self, name):                  # the single random_sample() inside np.random.choice(p=)
        return self._r(name, float(np.random.random_sample()))

    def gauss(self, shape, name):               # legacy_gauss draws in C order (normal = loc + scale*gauss)
        return self._r(name, np.random.standard_normal(shape))

    def randint(self, n, size, name):           # envs/darkroom_env.py:24,27
        return self._r(name, np.random.randint(0, n, size))

    def uniform_vec(self, d, name):             # envs/bandit_env.py:12
        return self._r(name, np.random.uniform(0, 1, d))

    def arrays(self):
        return {k: np.stack(v) for k, v in self.rec.items()}


class ReplayNoise:
    """Replays named draw sequences (arrays indexed by call count)."""

    def __init__(self, arrays):
        self.a = {k: np.asarray(v) for k, v in arrays.items()}
        self.i = {k: 0 for k in arrays}

    def _n(self, name):
        v = self.a[name][self.i[name]]
        self.i[name] += 1
        return v

    def choice_index(self, n, name):
        return int(self._n(name))

    def choice_index_vec(self, n, size, name):
        return np.asarray(self._n(name)).astype(np.int64)

    def dirichlet(self, d, name):
        return np.asarray(self._n(name), dtype=np.float64)

    def uniform01(self, name):
        return float(self._n(name))

    def gauss(self, shape, name):
        v = np.asarray(self._n(name), dtype=np.float64)
        return v if shape is not None else float(v)

    def randint(self, n, size, name):
        return np.asarray(self._n(name)).astype(np.int64)

    def uniform_vec(self, d, name):
        return np.asarray(self._n(name), dtype=np.float64)


# ----------------------------------------------------------------------------
# bandit task + rollin_bandit  (configs 1 / 5)
# ----------------------------------------------------------------------------
def sample_bandit_means(n_envs, dim, noise):
    """envs/bandit_env.py:10-18 via collect_data.py:222 -- all env draws come first."""
    return np.stack([noise.uniform_vec(dim, "means") for _ in range(n_envs)])


def opt_action(means):
    """envs/bandit_env.py:30-34: first argmax, one-hot."""
    a = np.zeros(means.shape)
    a[np.argmax(means)] = 1.0
    return a


def choice_cdf(probs):
    """numpy legacy RandomState.choice(p=): cdf = cumsum(p); cdf /= cdf[-1]."""
    cdf = np.cumsum(probs)
    cdf /= cdf[-1]
    return cdf


def rollin_bandit(means, H, var, noise):
    """collect_data.py:23-53 (behaviour policy) + envs/bandit_env.py:56-64 (transit).

    Returns xs (H,1) int64, us (H,d) f64 one-hot, xps (H,1) int64, rs (H,) f64.
    ``cov`` argument of the reference is overwritten at :30, so it is not a parameter."""
    d = len(means)
    cov = COV_GRID[noise.choice_index(11, "cov_idx")]           # :30
    probs = noise.dirichlet(d, "dir_probs")                     # :31-32
    probs2 = np.zeros(d)
    probs2[noise.choice_index(d, "rand_idx")] = 1.0             # :33-35
    probs = (1 - cov) * probs + cov * probs2                    # :36
    cdf = choice_cdf(probs)
    xs = np.ones((H, 1), dtype=np.int64)
    us = np.zeros((H, d))
    rs = np.zeros(H)
    for h in range(H):                                          # :40
        i = int(cdf.searchsorted(noise.uniform01("u"), side="right"))   # :43
        us[h, i] = 1.0
        rs[h] = means[i] + (0.0 + var * noise.gauss(None, "z"))        # bandit_env.py:59
    return xs, us, xs.copy(), rs


def rollin_bandit_batch(means, var, cov_idx, dir_probs, rand_idx, u, z, reward_type="uniform"):
    """Vectorised form of ``rollin_bandit`` over N envs with explicit noise arrays
    (same float64 operations per element, so bit-identical to the loop form).
    means [N,d], cov_idx [N], dir_probs [N,d], rand_idx [N], u [N,H], z [N,H].
    reward_type 'bernoulli' (envs/bandit_env.py:60-61): r ~ Bernoulli(means[a]) drawn as [z < means[a]] with z
    a uniform in [0,1) -- torch.bernoulli's rule (envs/gpu_bandit_env.py:60); np.random.binomial maps its uniform
    differently, so for this type parity with the reference is distributional, exact only on the device's own u."""
    means = np.asarray(means, dtype=np.float64)
    N, d = means.shape
    cov = np.asarray(COV_GRID)[np.asarray(cov_idx)][:, None]
    probs2 = np.zeros((N, d))
    probs2[np.arange(N), np.asarray(rand_idx)] = 1.0
    probs = (1 - cov) * np.asarray(dir_probs, dtype=np.float64) + cov * probs2
    cdf = np.cumsum(probs, axis=1)           # sequential adds in index order, like 1-D cumsum
    cdf = cdf / cdf[:, -1:]
    u = np.asarray(u, dtype=np.float64)
    acts = (cdf[:, None, :] <= u[:, :, None]).sum(-1)          # searchsorted(side='right')
    us = np.zeros((N, u.shape[1], d))
    np.put_along_axis(us, acts[:, :, None], 1.0, axis=2)
    ma = np.take_along_axis(means, acts, axis=1)
    if reward_type == "bernoulli":
        rs = (np.asarray(z, dtype=np.float64) < ma).astype(np.float64)
    else:
        rs = ma + (0.0 + var * np.asarray(z, dtype=np.float64))
    xs = np.ones((N, u.shape[1], 1), dtype=np.int64)
    return xs, us, xs.copy(), rs, acts


def generate_bandit_histories(n_envs, dim, horizon, var, noise, n_hists=1, n_samples=1):
    """collect_data.py:221-225 + :158-182: list of traj dicts (same keys / dtypes)."""
    all_means = sample_bandit_means(n_envs, dim, noise)
    trajs = []
    for means in all_means:
        for _ in range(n_hists):
            xs, us, xps, rs = rollin_bandit(means, horizon, var, noise)
            for _ in range(n_samples):
                trajs.append({
                    "query_state": np.array([1]), "optimal_action": opt_action(means),
                    "context_states": xs, "context_actions": us,
                    "context_next_states": xps, "context_rewards": rs, "means": means,
                })
    return trajs


# ----------------------------------------------------------------------------
# darkroom + rollin_mdp  (config 2)
# ----------------------------------------------------------------------------
def darkroom_transit(state, a_idx, goal, dim, perm=None):
    """envs/darkroom_env.py:37-55 (+ :100-103 for the permuted variant)."""
    if perm is not None:
        a_idx = perm[a_idx]
    s = np.array(state, dtype=np.int64)
    if a_idx == 0:
        s[0] += 1
    elif a_idx == 1:
        s[0] -= 1
    elif a_idx == 2:
        s[1] += 1
    elif a_idx == 3:
        s[1] -= 1
    s = np.clip(s, 0, dim - 1)
    return s, int(np.all(s == np.asarray(goal)))


def darkroom_opt_action_index(state, goal, perm=None):
    """envs/darkroom_env.py:69-82 (+ :105-111: index j with perm[j] == base action)."""
    if state[0] < goal[0]:
        a = 0
    elif state[0] > goal[0]:
        a = 1
    elif state[1] < goal[1]:
        a = 2
    elif state[1] > goal[1]:
        a = 3
    else:
        a = 4
    if perm is not None:
        a = list(perm).index(a)
    return a


def rollin_mdp(goal, dim, H, rollin_type, noise, perm_index=None):
    """collect_data.py:83-111.  Returns states (H,2) i64, actions (H,5) f64, next_states (H,2) i64, rewards (H,) i64."""
    perm = None if perm_index is None else DARKROOM_PERMS[perm_index]
    states = np.zeros((H, 2), dtype=np.int64)
    actions = np.zeros((H, 5))
    next_states = np.zeros((H, 2), dtype=np.int64)
    rewards = np.zeros(H, dtype=np.int64)
    state = np.array([0, 0])                                        # reset, darkroom_env.py:32-35
    for h in range(H):
        if rollin_type == "uniform":
            state = noise.randint(dim, 2, "state")                  # :92 -> darkroom_env.py:24
            a = int(noise.randint(5, None, "action"))               # :93 -> darkroom_env.py:27
        elif rollin_type == "expert":
            a = darkroom_opt_action_index(state, goal, perm)        # :95
        else:
            raise NotImplementedError
        ns, r = darkroom_transit(state, a, goal, dim, perm)         # :98
        states[h], next_states[h], rewards[h] = state, ns, r
        actions[h, a] = 1.0
        state = ns                                                  # :104
    return states, actions, next_states, rewards


def generate_mdp_histories(goals, dim, H, rollin_type, noise, perm_indices=None, n_hists=1, n_samples=1):
    """collect_data.py:189-218 / :290-300."""
    trajs = []
    for e, goal in enumerate(goals):
        pi = None if perm_indices is None else int(perm_indices[e])
        perm = None if pi is None else DARKROOM_PERMS[pi]
        for _ in range(n_hists):
            s, a, ns, r = rollin_mdp(goal, dim, H, rollin_type, noise, pi)
            for _ in range(n_samples):
                q = noise.randint(dim, 2, "query")                  # :200
                oa = np.zeros(5)
                oa[darkroom_opt_action_index(q, goal, perm)] = 1.0  # :201
                t = {"query_state": q, "optimal_action": oa, "context_states": s, "context_actions": a,
                     "context_next_states": ns, "context_rewards": r, "goal": np.array(goal)}
                if pi is not None:
                    t["perm_index"] = pi
                trajs.append(t)
    return trajs


def darkroom_transit_batch(states, a_idx, goals, dim, perms=None):
    """Vectorised darkroom_transit: states [...,2] int, a_idx [...], goals broadcastable [...,2]."""
    states = np.asarray(states, dtype=np.int64)
    a = np.asarray(a_idx, dtype=np.int64)
    if perms is not None:
        a = np.take_along_axis(np.broadcast_to(perms, a.shape + (5,)), a[..., None], -1)[..., 0]
    dx = (a == 0).astype(np.int64) - (a == 1)
    dy = (a == 2).astype(np.int64) - (a == 3)
    ns = np.stack([states[..., 0] + dx, states[..., 1] + dy], -1)
    ns = np.clip(ns, 0, dim - 1)
    r = np.all(ns == np.asarray(goals), axis=-1).astype(np.int64)
    return ns, r


# ----------------------------------------------------------------------------
# controllers + deploy_online_vec  (configs 3 / 4)
# ----------------------------------------------------------------------------
def _arm_stats(actions, rewards, d):
    """Per-(env, arm) reward sums and counts recomputed from the context, as every
    classical controller does each step (ctrls/ctrl_bandit.py:95-104, :355-364)."""
    N = actions.shape[0]
    b = np.zeros((N, d))
    counts = np.zeros((N, d))
    idx = np.argmax(actions, axis=-1) if actions.shape[1] else np.zeros((N, 0), dtype=np.int64)
    for e in range(N):
        for c in range(d):
            ar = rewards[e][idx[e] == c]
            b[e, c] = np.sum(ar)
            counts[e, c] = len(ar)
    return b, counts


class OptCtrl:
    """ctrls/ctrl_bandit.py:22-38."""
    def __init__(self, means):
        self.opt = np.stack([opt_action(m) for m in means])

    def act(self, ctx_a, ctx_r, noise, h):
        return self.opt.copy()


class EmpMeanCtrl:
    """ctrls/ctrl_bandit.py:57-118 (act_numpy_vec :91-118)."""
    def __init__(self, d, online=False):
        self.d, self.online = d, online

    def act(self, ctx_a, ctx_r, noise, h):
        N = ctx_a.shape[0]
        b, counts = _arm_stats(ctx_a, ctx_r, self.d)
        b_mean = b / np.maximum(1, counts)
        i = np.argmax(b_mean, axis=-1)
        j = np.argmin(counts, axis=-1)
        if self.online:
            mask = counts[np.arange(N), j] == 0
            i[mask] = j[mask]
        a = np.zeros((N, self.d))
        a[np.arange(N), i] = 1.0
        return a


class UCBCtrl:
    """ctrls/ctrl_bandit.py:318-380.  The reference hard-codes the untried-arm mask to 200
    envs (:374, an IndexError for any other batch size); this restatement uses the batch size,
    which is identical at N=200 (the only size the reference can run)."""
    def __init__(self, d, const=1.0):
        self.d, self.const = d, const

    def act(self, ctx_a, ctx_r, noise, h):
        N = ctx_a.shape[0]
        b, counts = _arm_stats(ctx_a, ctx_r, self.d)
        b_mean = b / np.maximum(1, counts)
        bounds = b_mean + self.const / np.maximum(1, np.sqrt(counts))
        i = np.argmax(bounds, axis=-1)
        j = np.argmin(counts, axis=-1)
        mask = counts[np.arange(N), j] == 0
        i[mask] = j[mask]
        a = np.zeros((N, self.d))
        a[np.arange(N), i] = 1.0
        return a


class ThompsonCtrl:
    """ctrls/ctrl_bandit.py:122-251 with batch_size>1, sample=True path
    (set_batch_numpy_vec :159-182, update_posterior_all :196-203, act_numpy_vec :232-251)."""
    def __init__(self, d, std=.1, sample=True, prior_mean=.5, prior_var=1 / 12.0, n_mode_draws=100):
        self.d, self.variance = d, std ** 2
        self.prior_mean, self.prior_variance = prior_mean, prior_var
        self.sample, self.n_mode_draws = sample, n_mode_draws

    def posterior(self, ctx_a, ctx_r):
        N = ctx_a.shape[0]
        means = np.ones((N, self.d)) * self.prior_mean
        variances = np.ones((N, self.d)) * self.prior_variance
        b, counts = _arm_stats(ctx_a, ctx_r, self.d)
        arm_means = np.zeros((N, self.d))
        idx = np.argmax(ctx_a, axis=-1) if ctx_a.shape[1] else np.zeros((N, 0), dtype=np.int64)
        for e in range(N):
            for c in range(self.d):
                if counts[e, c] > 0:
                    arm_means[e, c] = np.mean(ctx_r[e][idx[e] == c])       # :175-177
        with np.errstate(divide="ignore", invalid="ignore"):
            prior_weight = self.variance / (self.variance + counts * self.prior_variance)
            new_mean = prior_weight * self.prior_mean + (1 - prior_weight) * arm_means
            new_var = 1 / (1 / self.prior_variance + counts / self.variance)
        mask = counts > 0
        means[mask] = new_mean[mask]
        variances[mask] = new_var[mask]
        return means, variances

    def act(self, ctx_a, ctx_r, noise, h):
        N = ctx_a.shape[0]
        means, variances = self.posterior(ctx_a, ctx_r)
        if self.sample:
            values = means + np.sqrt(variances) * noise.gauss((N, self.d), "thompson_z")   # :234
            i = np.argmax(values, axis=-1)
        else:
            values = np.stack([means + np.sqrt(variances) * noise.gauss((N, self.d), "thompson_z")
                               for _ in range(self.n_mode_draws)], axis=1)                 # :241-244
            amax = np.argmax(values, axis=-1)
            freqs = np.array([np.bincount(am, minlength=self.d) for am in amax])
            i = np.argmax(freqs, axis=-1)
        a = np.zeros((N, self.d))
        a[np.arange(N), i] = 1.0
        return a


class LinUCBCtrl:
    """ctrls/ctrl_bandit.py:447-528 (act_numpy_vec :491-528)."""
    def __init__(self, arms, const=1.0):
        self.arms, self.const = np.asarray(arms, dtype=np.float64), const
        self.dim, self.lin_d = self.arms.shape

    def act(self, ctx_a, ctx_r, noise, h):
        N = ctx_a.shape[0]
        hot = np.zeros((N, self.dim))
        if ctx_r.shape[1] < 1:
            hot[np.arange(N), noise.choice_index_vec(self.dim, N, "linucb_first")] = 1     # :496-500
            return hot
        for e in range(N):
            A = self.arms[np.argmax(ctx_a[e], axis=1)]
            cov = np.eye(self.lin_d) + A.T @ A
            cov_inv = np.linalg.inv(cov)
            theta = (cov_inv @ A.T @ ctx_r[e]).flatten()
            best, best_v = None, -np.inf
            for i, arm in enumerate(self.arms):
                v = theta @ arm + self.const * np.sqrt(arm @ cov_inv @ arm)
                if v > best_v:
                    best_v, best = v, i
            hot[e, best] = 1
        return hot


class TransformerCtrl:
    """ctrls/ctrl_bandit.py:383-444: logits = model(batch)[:, -1]; sample=True ->
    scipy softmax + np.random.choice(p) per env (:436-438), else argmax (:440).
    ``logits_fn(ctx_s, ctx_a, ctx_ns, ctx_r) -> [N,d]`` float array."""
    def __init__(self, logits_fn, d, sample=True):
        self.f, self.d, self.sample = logits_fn, d, sample

    def act_full(self, ctx, noise, h):
        logits = np.asarray(self.f(*ctx), dtype=np.float64)
        N = logits.shape[0]
        if self.sample:
            m = logits.max(axis=-1, keepdims=True)          # scipy.special.softmax
            e = np.exp(logits - m)
            probs = e / e.sum(axis=-1, keepdims=True)
            i = np.array([int(choice_cdf(p).searchsorted(noise.uniform01("ctrl_u"), side="right")) for p in probs])
        else:
            i = np.argmax(logits, axis=-1)
        a = np.zeros((N, self.d))
        a[np.arange(N), i] = 1.0
        return a


def deploy_online_vec(means, var, H, ctrl, noise, include_meta=True, reward_type="uniform"):
    """evals/eval_bandit.py:56-103 driving envs/bandit_env.py:125-149 (BanditEnvVec.deploy,
    one step per call because BanditEnv.H == 1) and :98-105/:56-64 (step/transit).

    Returns cum_means [H,N] f64 and the four context arrays [N,H,.] f64."""
    means = np.asarray(means, dtype=np.float64)
    N, d = means.shape
    ctx_s = np.zeros((N, H, 1))
    ctx_a = np.zeros((N, H, d))
    ctx_ns = np.zeros((N, H, 1))
    ctx_r = np.zeros((N, H, 1))
    cum = []
    for h in range(H):
        if isinstance(ctrl, TransformerCtrl):
            u = ctrl.act_full((ctx_s[:, :h], ctx_a[:, :h], ctx_ns[:, :h], ctx_r[:, :h]), noise, h)
        else:
            u = ctrl.act(ctx_a[:, :h], ctx_r[:, :h, 0], noise, h)
        a = np.argmax(u, axis=-1)
        z = noise.gauss((N,), "reward_z")                      # N x np.random.normal(0, var) in env order
        if reward_type == "bernoulli":                          # bandit_env.py:60-61; z is then a U[0,1) draw
            r = (z < means[np.arange(N), a]).astype(np.float64)
        else:
            r = means[np.arange(N), a] + (0.0 + var * z)
        ctx_s[:, h, 0] = 1
        ctx_a[:, h] = u
        ctx_ns[:, h, 0] = 1
        ctx_r[:, h, 0] = r
        cum.append(np.sum(means * u, axis=-1))                 # get_arm_value, bandit_env.py:151-153
    return np.array(cum), {"context_states": ctx_s, "context_actions": ctx_a,
                           "context_next_states": ctx_ns, "context_rewards": ctx_r}


def deploy_online_vec_darkroom(goals, dim, Heps, H, horizon, logits_fn, noise, perm_indices=None, sample=True):
    """evals/eval_darkroom.py:20-84 driving DarkroomEnvVec.deploy (envs/darkroom_env.py:151-175) and
    DarkroomTransformerController.act (ctrls/ctrl_darkroom.py:35-66).  ``logits_fn(query [N,2], cs, ca,
    cns, cr) -> [N,5]``.  Returns per-episode returns [N, Heps] and the last context."""
    assert H % horizon == 0
    ctx_rollouts = H // horizon
    N = len(goals)
    goals = np.asarray(goals)
    perms = None if perm_indices is None else [DARKROOM_PERMS[int(p)] for p in perm_indices]
    cs = np.zeros((N, ctx_rollouts, horizon, 2))
    ca = np.zeros((N, ctx_rollouts, horizon, 5))
    cns = np.zeros((N, ctx_rollouts, horizon, 2))
    cr = np.zeros((N, ctx_rollouts, horizon, 1))
    cum = []
    for ep in range(Heps):
        k = min(ep, ctx_rollouts)
        ctx = (cs[:, :k].reshape(N, -1, 2), ca[:, :k].reshape(N, -1, 5), cns[:, :k].reshape(N, -1, 2), cr[:, :k].reshape(N, -1, 1))
        state = np.zeros((N, 2), dtype=np.int64)                                   # reset, darkroom_env.py:32-35
        S, A, NS, R = (np.zeros((N, horizon, 2)), np.zeros((N, horizon, 5)), np.zeros((N, horizon, 2)), np.zeros((N, horizon)))
        for t in range(horizon):
            logits = np.asarray(logits_fn(state.astype(np.float64), *ctx), dtype=np.float64)
            if sample:
                e = np.exp(logits - logits.max(-1, keepdims=True))                # scipy softmax, temp = 1
                probs = e / e.sum(-1, keepdims=True)
                a = np.array([int(choice_cdf(p).searchsorted(noise.uniform01("ctrl_u"), side="right")) for p in probs])
            else:
                a = logits.argmax(-1)
            for e_ in range(N):
                ns, r = darkroom_transit(state[e_], int(a[e_]), goals[e_], dim, None if perms is None else perms[e_])
                S[e_, t], NS[e_, t], R[e_, t] = state[e_], ns, r
                A[e_, t, a[e_]] = 1.0
                state[e_] = ns
        cum.append(R.sum(-1))
        if ep < ctx_rollouts:
            cs[:, ep], ca[:, ep], cns[:, ep], cr[:, ep, :, 0] = S, A, NS, R
        else:
            cs = np.concatenate([cs[:, 1:], S[:, None]], 1)
            ca = np.concatenate([ca[:, 1:], A[:, None]], 1)
            cns = np.concatenate([cns[:, 1:], NS[:, None]], 1)
            cr = np.concatenate([cr[:, 1:], R[:, None, :, None]], 1)
    return np.stack(cum, axis=1), (cs, ca, cns, cr)


def darkroom_episode(goals, dim, horizon, logits_fn, u, perm_indices=None, sample=True):
    """One DarkroomEnvVec.deploy episode (envs/darkroom_env.py:151-175) under DarkroomTransformerController.act
    (ctrls/ctrl_darkroom.py:35-66) with a FIXED context: ``logits_fn(query [N,2]) -> [N,5]``; ``u`` [horizon,N]
    are the uniforms of the categorical draws (env order within a step).  Returns the per-env returns [N] --
    the learner half of evals/eval_darkroom.py:124-190 (`offline`)."""
    goals = np.asarray(goals)
    N = len(goals)
    perms = None if perm_indices is None else [DARKROOM_PERMS[int(p)] for p in perm_indices]
    state = np.zeros((N, 2), dtype=np.int64)
    ret = np.zeros(N)
    for t in range(horizon):
        logits = np.asarray(logits_fn(state.astype(np.float64)), dtype=np.float64)
        if sample:
            e = np.exp(logits - logits.max(-1, keepdims=True))
            probs = e / e.sum(-1, keepdims=True)
            a = np.array([int(choice_cdf(p).searchsorted(u[t][i], side="right")) for i, p in enumerate(probs)])
        else:
            a = logits.argmax(-1)
        for i in range(N):
            ns, r = darkroom_transit(state[i], int(a[i]), goals[i], dim, None if perms is None else perms[i])
            state[i] = ns
            ret[i] += r
    return ret


def regret_stats(opt_means, alg_means):
    """evals/eval_bandit.py:169-178: inputs [N,H]; returns per-step mean, sem and cumulative mean, sem."""
    diff = np.asarray(opt_means) - np.asarray(alg_means)
    n = diff.shape[0]

    def sem(v):                                                # scipy.stats.sem, ddof=1
        return np.std(v, axis=0, ddof=1) / np.sqrt(n)
    cr = np.cumsum(diff, axis=1)
    return diff.mean(0), sem(diff), cr.mean(0), sem(cr)


def linear_bandit_arms(dim, lin_d):
    """collect_data.py:230-231: fixed arm features shared by all envs."""
    rng = np.random.RandomState(seed=1234)
    return rng.normal(size=(dim, lin_d)) / np.sqrt(lin_d)


def sample_linear_thetas(n_envs, lin_d, noise):
    """envs/bandit_env.py:21-25 via collect_data.py:233."""
    return np.stack([noise.gauss((lin_d,), "theta") / np.sqrt(lin_d) for _ in range(n_envs)])


# ----------------------------------------------------------------------------
# GPT-2 trunk restatement (models/net.py:41-60 over transformers GPT2Model, n_head=1)
# ----------------------------------------------------------------------------
def _ln(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def _gelu_new(x):
    """transformers/activations.py NewGELUActivation (tanh form)."""
    return 0.5 * x * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * x ** 3)))


def transformer_forward(sd, query_states, ctx_s, ctx_a, ctx_ns, ctx_r, n_layer, test=True, dtype=np.float64):
    """models/net.py:41-60.  ``sd``: dict of numpy arrays with the reference state_dict keys
    (transformer.wpe.weight, transformer.h.{l}.ln_1.weight ... embed_transition.*, pred_actions.*).
    Query token at position 0 (zero action / next_state / reward), context after it; pre-LN
    GPT-2 blocks with one head (head_dim = n_embd), causal softmax(QK^T/sqrt(n_embd)),
    gelu_new MLP, final ln_f, linear head.  Returns [B,du] if test else [B,T,du]."""
    f = lambda k: np.asarray(sd[k], dtype=dtype)
    B, T = ctx_a.shape[0], ctx_a.shape[1]
    dx, du = query_states.shape[-1], ctx_a.shape[-1]
    seq = np.zeros((B, T + 1, 2 * dx + du + 1), dtype=dtype)
    seq[:, 0, :dx] = query_states
    seq[:, 1:, :dx] = ctx_s
    seq[:, 1:, dx:dx + du] = ctx_a
    seq[:, 1:, dx + du:2 * dx + du] = ctx_ns
    seq[:, 1:, -1:] = np.asarray(ctx_r).reshape(B, T, 1)
    x = seq @ f("embed_transition.weight").T + f("embed_transition.bias")
    E = x.shape[-1]
    x = x + f("transformer.wpe.weight")[: T + 1]
    mask = np.tril(np.ones((T + 1, T + 1), dtype=bool))
    for l in range(n_layer):
        p = "transformer.h.%d." % l
        h = _ln(x, f(p + "ln_1.weight"), f(p + "ln_1.bias"))
        qkv = h @ f(p + "attn.c_attn.weight") + f(p + "attn.c_attn.bias")
        q, k, v = qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:]
        s = q @ np.swapaxes(k, -1, -2) / np.sqrt(dtype(E))
        s = np.where(mask, s, -np.inf)
        s = s - s.max(-1, keepdims=True)
        w = np.exp(s)
        w = w / w.sum(-1, keepdims=True)
        a = w @ v
        x = x + a @ f(p + "attn.c_proj.weight") + f(p + "attn.c_proj.bias")
        h = _ln(x, f(p + "ln_2.weight"), f(p + "ln_2.bias"))
        m = _gelu_new(h @ f(p + "mlp.c_fc.weight") + f(p + "mlp.c_fc.bias"))
        x = x + m @ f(p + "mlp.c_proj.weight") + f(p + "mlp.c_proj.bias")
    x = _ln(x, f("transformer.ln_f.weight"), f("transformer.ln_f.bias"))
    preds = x @ f("pred_actions.weight").T + f("pred_actions.bias")
    return preds[:, -1, :] if test else preds[:, 1:, :]
