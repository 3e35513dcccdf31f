"""CPU tier: the oracle against the golden vectors produced from the live reference
(oracle/make_golden.py), the Philox known-answer vectors, and the host-side logic."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import dpt_oracle as O
from oracle import philox as P


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32_10
    kat = [([0, 0, 0, 0], (0, 0), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, (0xffffffff, 0xffffffff), [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0),
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert [int(x) for x in P.philox4x32_10(np.array(ctr, dtype=np.uint64), key)] == want


def test_philox_bounded_draws_in_range():
    st, a = P.darkroom_draws(3, np.arange(50), 37, 10)
    assert st.min() >= 0 and st.max() <= 9 and a.min() >= 0 and a.max() <= 4
    assert len(np.unique(st[..., 0])) == 10 and len(np.unique(a)) == 5
    m = P.bandit_means(0, np.arange(1000), 5)
    assert m.min() >= 0 and m.max() < 1 and abs(m.mean() - 0.5) < 0.02
    assert np.array_equal(m.astype(np.float32).astype(np.float64), m)   # fp32-exact


@pytest.mark.parametrize("name", ["bandit_rollin_d5", "bandit_rollin_d10", "bandit_rollin_d3"])
def test_bandit_rollin_golden(name):
    g = golden(name)
    n, H = g["u"].shape
    # loop form through ReplayNoise
    noise = O.ReplayNoise({"cov_idx": g["cov_idx"], "dir_probs": g["dir_probs"], "rand_idx": g["rand_idx"],
                           "u": g["u"].reshape(-1), "z": g["z"].reshape(-1)})
    for e in range(n):
        xs, us, xps, rs = O.rollin_bandit(g["means"][e], H, float(g["var"]), noise)
        assert np.array_equal(us, g["ref_actions"][e]) and np.array_equal(rs, g["ref_rewards"][e])
        assert np.array_equal(xs, g["ref_states"][e]) and xs.dtype == np.int64
        assert np.array_equal(O.opt_action(g["means"][e]), g["ref_optimal_action"][e])
    # vectorised form
    xs, us, xps, rs, acts = O.rollin_bandit_batch(g["means"], float(g["var"]), g["cov_idx"], g["dir_probs"],
                                                  g["rand_idx"], g["u"], g["z"])
    assert np.array_equal(us, g["ref_actions"]) and np.array_equal(rs, g["ref_rewards"])


@pytest.mark.parametrize("name", ["darkroom_uniform", "darkroom_expert", "darkroom_perm_uniform", "darkroom_perm_expert"])
def test_darkroom_golden(name):
    g = golden(name)
    goals, H, dim = g["goals"], int(g["H"]), int(g["dim"])
    perms = g["perm_indices"] if len(g["perm_indices"]) else None
    mode = str(g["rollin_type"])
    arrays = {"query": g["query"]}
    if mode == "uniform":
        arrays.update(state=g["state"].reshape(-1, 2), action=g["action"].reshape(-1))
    # per-env replay: noise order is env-major (rollin then query)
    for e in range(len(goals)):
        a = {"query": g["query"][e:e + 1]}
        if mode == "uniform":
            a.update(state=g["state"][e], action=g["action"][e])
        t = O.generate_mdp_histories([goals[e]], dim, H, mode, O.ReplayNoise(a),
                                     None if perms is None else [perms[e]])[0]
        assert np.array_equal(t["context_states"], g["ref_states"][e])
        assert np.array_equal(t["context_actions"].argmax(-1), g["ref_actions"][e])
        assert np.array_equal(t["context_next_states"], g["ref_next_states"][e])
        assert np.array_equal(t["context_rewards"], g["ref_rewards"][e])
        assert t["optimal_action"].argmax() == g["ref_optimal_action"][e]
    if mode == "uniform":   # vectorised transit
        pt = None if perms is None else np.asarray(O.DARKROOM_PERMS)[perms][:, None, :]
        ns, r = O.darkroom_transit_batch(g["state"], g["action"], goals[:, None, :], dim, pt)
        assert np.array_equal(ns, g["ref_next_states"]) and np.array_equal(r, g["ref_rewards"])


def test_darkroom_exhaustive_table():
    g = golden("darkroom_table")
    dim = int(g["dim"])
    assert np.array_equal(g["perm_table"], np.asarray(O.DARKROOM_PERMS))
    xs, ys, acts = np.meshgrid(np.arange(dim), np.arange(dim), np.arange(5), indexing="ij")
    st = np.stack([xs, ys], -1)
    for i, goal in enumerate(g["goals"]):
        ns, r = O.darkroom_transit_batch(st, acts, goal, dim)
        assert np.array_equal(ns, g["next_state"][i]) and np.array_equal(r, g["reward"][i])
    for i, pi in enumerate(g["perms"]):
        ns, r = O.darkroom_transit_batch(st, acts, [dim - 1, dim - 1], dim, np.asarray(O.DARKROOM_PERMS[pi]))
        assert np.array_equal(ns, g["p_next_state"][i]) and np.array_equal(r, g["p_reward"][i])


@pytest.mark.parametrize("name,ctrls", [("online_d5_n200", ["opt", "emp", "emp_offline", "ucb", "thompson"]),
                                        ("online_d10_n16", ["opt", "emp", "thompson"])])
def test_online_golden(name, ctrls):
    g = golden(name)
    means, H, var = g["means"], int(g["H"]), float(g["var"])
    N, d = means.shape
    mk = {"opt": lambda: O.OptCtrl(means), "emp": lambda: O.EmpMeanCtrl(d, online=True),
          "emp_offline": lambda: O.EmpMeanCtrl(d, online=False), "ucb": lambda: O.UCBCtrl(d, 1.0),
          "thompson": lambda: O.ThompsonCtrl(d, std=var, sample=True, prior_mean=.5, prior_var=1 / 12.0)}
    for c in ctrls:
        arrays = {"reward_z": g[c + "_reward_z"]}
        if c == "thompson":
            arrays["thompson_z"] = g[c + "_thompson_z"]
        cum, meta = O.deploy_online_vec(means, var, H, mk[c](), O.ReplayNoise(arrays))
        assert np.array_equal(meta["context_actions"].argmax(-1), g[c + "_actions"]), c
        assert np.array_equal(meta["context_rewards"][:, :, 0], g[c + "_rewards"]), c
        assert np.array_equal(cum, g[c + "_cum_means"]), c


def test_linear_bandit_golden():
    g = golden("linear_bandit")
    arms, means, H, var = g["arms"], g["means"], int(g["H"]), float(g["var"])
    assert np.array_equal(arms, O.linear_bandit_arms(*arms.shape))
    assert np.array_equal(means, np.stack([arms @ t for t in g["thetas"]]))
    cum, meta = O.deploy_online_vec(means, var, H, O.ThompsonCtrl(arms.shape[0], std=var, sample=True, prior_mean=0.0, prior_var=1.0),
                                    O.ReplayNoise({"reward_z": g["thompson_reward_z"], "thompson_z": g["thompson_thompson_z"]}))
    assert np.array_equal(meta["context_actions"].argmax(-1), g["thompson_actions"])
    assert np.array_equal(meta["context_rewards"][:, :, 0], g["thompson_rewards"])
    cum, meta = O.deploy_online_vec(means, var, H, O.LinUCBCtrl(arms, 1.0),
                                    O.ReplayNoise({"reward_z": g["linucb_reward_z"], "linucb_first": g["linucb_first"][None]}))
    assert np.array_equal(meta["context_actions"].argmax(-1), g["linucb_actions"])
    assert np.array_equal(cum, g["linucb_cum_means"])


@pytest.mark.parametrize("name", ["transformer_l2", "transformer_l4"])
def test_transformer_forward_golden(name):
    g = golden(name)
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    H, L, d = int(g["H"]), int(g["n_layer"]), int(g["d"])
    acts, rew = g["fwd_actions"].astype(np.int64), g["fwd_rewards"]
    B = acts.shape[0]
    ca, cs, q = np.eye(d)[acts], np.ones((B, H, 1)), np.ones((B, 1))
    for t in (0, 1, 5, H):
        o = O.transformer_forward(sd, q, cs[:, :t], ca[:, :t], cs[:, :t], rew[:, :t], L, test=True)
        ref = g["logits_t%d" % t]
        assert np.abs(o - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
        if t > 0:
            o = O.transformer_forward(sd, q, cs[:, :t], ca[:, :t], cs[:, :t], rew[:, :t], L, test=False)
            ref = g["logits_all_t%d" % t]
            assert o.shape == ref.shape and np.abs(o - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_regret_stats_match_scipy():
    import scipy.stats
    rng = np.random.RandomState(0)
    opt, alg = rng.rand(50, 20), rng.rand(50, 20)
    m, s, cm, cs = O.regret_stats(opt, alg)
    diff = opt - alg
    assert np.allclose(m, diff.mean(0)) and np.allclose(s, scipy.stats.sem(diff, axis=0))
    assert np.allclose(cs, scipy.stats.sem(np.cumsum(diff, axis=1), axis=0))


@pytest.mark.parametrize("name", ["darkroom_online", "darkroom_online_perm"])
def test_darkroom_online_golden(name):
    """evals/eval_darkroom.py deploy_online_vec: oracle loop + oracle float64 forward on the recorded uniforms."""
    g = golden(name)
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    L = int(g["n_layer"])
    perms = g["perm_indices"] if len(g["perm_indices"]) else None

    def logits_fn(q, cs, ca, cns, cr):
        return O.transformer_forward(sd, q, cs, ca, cns, cr, L, test=True)
    ret, _ = O.deploy_online_vec_darkroom(g["goals"], int(g["dim"]), int(g["Heps"]), int(g["H"]), int(g["horizon"]), logits_fn,
                                          O.ReplayNoise({"ctrl_u": g["ctrl_u"].reshape(-1)}), perms)
    assert np.array_equal(ret, g["ref_returns"])


def test_dataset_dropin(tmp_path):
    """dataset.py:11-91 hand-off format: same items as the reference's Dataset (compared live when the
    reference is mounted, i.e. in the dev container), and the device-batch constructor agrees with it."""
    import pickle
    import torch
    import dpt_b200
    from dpt_b200.dataset import Dataset
    from oracle import ref_loader
    rs = np.random.RandomState(0)
    H, d = 12, 5
    trajs = [{"query_state": np.array([1]), "optimal_action": np.eye(d)[rs.randint(d)], "context_states": np.ones((H, 1), dtype=np.int64),
              "context_actions": np.eye(d)[rs.randint(0, d, H)], "context_next_states": np.ones((H, 1), dtype=np.int64),
              "context_rewards": rs.normal(size=H), "means": rs.rand(d)} for _ in range(7)]
    path = tmp_path / "trajs.pkl"
    pickle.dump(trajs, open(path, "wb"))
    cfg = {"shuffle": True, "horizon": H, "store_gpu": False, "state_dim": 1, "action_dim": d}
    ds = Dataset(str(path), cfg)
    assert len(ds) == 7
    torch.manual_seed(3)
    item = ds[2]
    assert set(item) == {"context_states", "context_actions", "context_next_states", "context_rewards", "query_states", "optimal_actions", "zeros"}
    assert item["context_rewards"].shape == (H, 1) and item["context_actions"].shape == (H, d) and item["zeros"].shape == (1 + d + 1,)
    assert all(v.dtype == torch.float32 for v in item.values())
    torch.manual_seed(3)
    perm = torch.randperm(H)
    assert torch.equal(item["context_rewards"][:, 0], torch.tensor(trajs[2]["context_rewards"]).float()[perm])
    assert torch.equal(item["context_actions"], torch.tensor(trajs[2]["context_actions"]).float()[perm])
    # from_batch (what collect_bandit returns, here on the CPU) == from_trajs
    batch = {"context_states": torch.ones(7, H, 1), "context_next_states": torch.ones(7, H, 1),
             "context_actions": torch.tensor(np.stack([t["context_actions"] for t in trajs])).float(),
             "context_rewards": torch.tensor(np.stack([t["context_rewards"] for t in trajs])).float()[:, :, None],
             "optimal_actions": torch.tensor(np.stack([t["optimal_action"] for t in trajs])).float()}
    cfg2 = dict(cfg, shuffle=False)
    a, b = Dataset.from_batch(batch, cfg2), Dataset.from_trajs(trajs, cfg2)
    for i in range(7):
        for k in a[i]:
            assert torch.equal(a[i][k], b[i][k]), k
    if ref_loader.available():
        ref = ref_loader.load()
        rds = ref.dataset.Dataset(str(path), cfg)
        for i in range(7):
            torch.manual_seed(i)
            x = rds[i]
            torch.manual_seed(i)
            y = ds[i]
            for k in x:
                assert torch.equal(x[k].cpu(), y[k]), k


def test_host_side_eval_helpers():
    """Host-only pieces of the eval callers: generate_eval_trajs (evals/eval_interactive_bandit.py:29-40) draws on
    np.random like the reference; ThompsonSamplingPolicy.update_posterior[_all] (ctrls/ctrl_bandit.py:184-203)."""
    import dpt_b200
    from dpt_b200.evals.eval_interactive_bandit import generate_eval_trajs
    from dpt_b200.ctrls.ctrl_bandit import ThompsonSamplingPolicy
    np.random.seed(3)
    t = generate_eval_trajs(4, 5, "uniform")
    np.random.seed(3)
    want = [np.random.uniform(0, 1, 5) for _ in range(4)]
    assert all(np.array_equal(a["means"], b) for a, b in zip(t, want))
    assert generate_eval_trajs(2, 3, "bernoulli")[0]["means"].shape == (3,)
    with pytest.raises(ValueError):
        generate_eval_trajs(1, 3, "poisson")

    class _Env:
        dim = 3
    p = ThompsonSamplingPolicy(_Env(), std=0.3, sample=True, prior_mean=0.5, prior_var=1 / 12.0, batch_size=2)
    p.counts = np.array([[0., 2., 5.], [1., 0., 0.]])
    arm_means = np.array([[0., .4, .9], [.2, 0., 0.]])
    p.update_posterior_all(arm_means)
    w = 0.09 / (0.09 + p.counts / 12.0)
    assert np.allclose(p.means, np.where(p.counts > 0, w * 0.5 + (1 - w) * arm_means, 0.5))
    assert np.allclose(p.variances, np.where(p.counts > 0, 1 / (12.0 + p.counts / 0.09), 1 / 12.0))
    q = ThompsonSamplingPolicy(_Env(), std=0.3, sample=True)
    q.counts = np.array([0., 3., 0.])
    q.update_posterior(1, np.array([.5, .7, .9]))
    q.update_posterior(0, np.array([]))
    assert np.isclose(q.means[1], (0.09 / (0.09 + 3 / 12.0)) * 0.5 + (1 - 0.09 / (0.09 + 3 / 12.0)) * 0.7) and q.means[0] == 0.5


def test_utils_filenames_match_reference():
    """utils.py:15-153 naming contract (dataset pickles / checkpoints).  Expected strings were produced by the
    reference's own utils.py in the dev container (see the commit that added this test)."""
    import dpt_b200
    from dpt_b200 import utils as U
    cfg = {"n_hists": 1, "n_samples": 2, "horizon": 500, "dim": 5, "var": 0.3, "cov": 0.0, "lin_d": 2, "rollin_type": "uniform",
           "shuffle": True, "lr": 0.001, "dropout": 0.0, "n_embd": 32, "n_layer": 4, "n_head": 1, "n_envs": 100000, "seed": 1}
    assert U.build_bandit_data_filename("bandit", 1000, cfg, 0) == "datasets/trajs_bandit_envs1000_hists1_samples2_H500_d5_var0.3_cov0.0_train.pkl"
    assert U.build_bandit_data_filename("bandit", 1000, cfg, 1) == "datasets/trajs_bandit_envs1000_hists1_samples2_H500_d5_var0.3_cov0.0_test.pkl"
    assert U.build_bandit_data_filename("bandit", 1000, cfg, 2) == "datasets/trajs_bandit_envs1000_H500_d5_var0.3_cov0.0_eval.pkl"
    assert U.build_linear_bandit_data_filename("bandit", 1000, cfg, 0) == "datasets/trajs_bandit_envs1000_hists1_samples2_H500_d5_lind2_var0.3_cov0.0_train.pkl"
    assert U.build_linear_bandit_data_filename("bandit", 1000, cfg, 2) == "datasets/trajs_bandit_envs1000_H500_d5_lind2_var0.3_cov0.0_eval.pkl"
    assert U.build_darkroom_data_filename("bandit", 1000, cfg, 1) == "datasets/trajs_bandit_envs1000_hists1_samples2_H500_d5_test.pkl"
    assert U.build_darkroom_data_filename("bandit", 1000, cfg, 2) == "datasets/trajs_bandit_envs1000_H500_d5_uniform_eval.pkl"
    assert U.build_bandit_model_filename("darkroom_heldout", cfg) == \
        "darkroom_heldout_shufTrue_lr0.001_do0.0_embd32_layer4_head1_envs100000_hists1_samples2_var0.3_cov0.0_H500_d5_seed1"
    assert U.build_linear_bandit_model_filename("darkroom_heldout", cfg) == \
        "darkroom_heldout_shufTrue_lr0.001_do0.0_embd32_layer4_head1_envs100000_hists1_samples2_var0.3_cov0.0_H500_d5_lind2_seed1"
    assert U.build_darkroom_model_filename("darkroom_heldout", cfg) == \
        "darkroom_heldout_shufTrue_lr0.001_do0.0_embd32_layer4_head1_envs100000_hists1_samples2_H500_d5_seed1"
    t = U.convert_to_tensor([[1, 2], [3, 4]], store_gpu=False)
    assert t.dtype == torch.float32 and t.shape == (2, 2)
