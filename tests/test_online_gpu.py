"""GPU tier: the fused online loop (dpt_online_loop) and dpt_arm_stats against the reference's
golden runs and against the oracle on identical noise.  Arms (integers) bit-exact, rewards and
expected rewards within 1e-5 relative."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import dpt_oracle as O

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def _close(a, b, rtol=1e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.all(np.abs(a - b) <= rtol * np.maximum(1.0, np.abs(b))), np.abs(a - b).max()


KW = {"opt": dict(kind="opt"), "emp": dict(kind="emp", p0=1.0), "emp_offline": dict(kind="emp", p0=0.0),
      "ucb": dict(kind="ucb", p0=1.0)}


@pytest.mark.parametrize("name,ctrls", [("online_d5_n200", ["opt", "emp", "emp_offline", "ucb", "thompson"]),
                                        ("online_d10_n16", ["opt", "emp", "thompson"])])
def test_online_loop_injected_matches_reference(dpt, name, ctrls):
    """Same reward / Thompson normals as the reference run -> the reference's arms, rewards, cum_means."""
    g = golden(name)
    means, H, var = g["means"], int(g["H"]), float(g["var"])
    for c in ctrls:
        kw = dict(KW.get(c, dict(kind="thompson", p0=var, p1=0.5, p2=1 / 12.0)))
        inj = {"reward_z": g[c + "_reward_z"]}
        if c == "thompson":
            inj["ctrl_z"] = g[c + "_thompson_z"]
        out = dpt.kernels.online_loop(kw.pop("kind"), torch.tensor(means, dtype=torch.float32), H, var, 0, inject=inj, **kw)
        assert np.array_equal(_np(out["context_actions"]).argmax(-1), g[c + "_actions"]), c
        assert np.array_equal(_np(out["context_actions"]).sum(-1), np.ones_like(g[c + "_rewards"])), c
        _close(_np(out["context_rewards"])[:, :, 0], g[c + "_rewards"])
        _close(_np(out["cum_means"]), g[c + "_cum_means"])
        assert bool((out["context_states"] == 1).all()) and bool((out["context_next_states"] == 1).all())
        opt = means.max(1)[None, :] - g[c + "_cum_means"]
        _close(_np(out["regret_sums"])[:, 0], opt.sum(1), 1e-6)
        _close(_np(out["regret_sums"])[:, 1], (opt ** 2).sum(1), 1e-6)
        cr = np.cumsum(opt, axis=0)
        _close(_np(out["regret_sums"])[:, 2], cr.sum(1), 1e-6)
        _close(_np(out["regret_sums"])[:, 3], (cr ** 2).sum(1), 1e-6)


def test_linear_bandit_injected_matches_reference(dpt):
    g = golden("linear_bandit")
    arms, means, H, var = g["arms"], g["means"], int(g["H"]), float(g["var"])
    m32 = torch.tensor(means, dtype=torch.float32)
    out = dpt.kernels.online_loop("thompson", m32, H, var, 0, p0=var, p1=0.0, p2=1.0,
                                  inject={"reward_z": g["thompson_reward_z"], "ctrl_z": g["thompson_thompson_z"]})
    # means = arms @ theta are not fp32-exact here: arms must still agree, values to fp32 precision
    assert np.array_equal(_np(out["context_actions"]).argmax(-1), g["thompson_actions"])
    _close(_np(out["context_rewards"])[:, :, 0], g["thompson_rewards"], 1e-5)
    out = dpt.kernels.online_loop("linucb", m32, H, var, 0, p0=1.0, arms=arms,
                                  inject={"reward_z": g["linucb_reward_z"], "first_arm": g["linucb_first"]})
    assert np.array_equal(_np(out["context_actions"]).argmax(-1), g["linucb_actions"])
    _close(_np(out["context_rewards"])[:, :, 0], g["linucb_rewards"], 1e-5)
    _close(_np(out["cum_means"]), g["linucb_cum_means"], 1e-5)


@pytest.mark.parametrize("kind,N,d,H", [("opt", 70, 5, 33), ("emp", 300, 5, 60), ("ucb", 300, 5, 60), ("thompson", 200, 5, 64),
                                        ("emp", 40, 10, 37), ("ucb", 33, 16, 30), ("thompson", 50, 20, 21),
                                        ("linucb", 120, 10, 40), ("linucb", 40, 7, 25)])
def test_online_loop_philox_matches_oracle(dpt, kind, N, d, H):
    """Philox mode: dump the noise the device used, replay it through the oracle's faithful
    (recount-from-context) controllers: arms bit-exact, rewards / cum_means within 1e-5."""
    seed, env_id0, var = 77, 4000, 0.3
    means, _, _ = dpt.kernels.bandit_sample_means(N, d, seed, env_id0)
    m64 = _np(means).astype(np.float64)
    arms = None
    if kind == "linucb":
        lin_d = 2 if d == 10 else 3
        arms = O.linear_bandit_arms(d, lin_d)
        m64 = (arms @ (np.random.RandomState(3).normal(size=(lin_d, N)) / np.sqrt(lin_d))).T.astype(np.float32).astype(np.float64)
        means = torch.tensor(m64, dtype=torch.float32)
    par = {"opt": {}, "emp": dict(p0=1.0), "ucb": dict(p0=1.0), "thompson": dict(p0=var, p1=0.5, p2=1 / 12.0),
           "linucb": dict(p0=1.0, arms=arms)}[kind]
    out = dpt.kernels.online_loop(kind, means, H, var, seed, env_id0, dump=True, **par)
    nz = {k: _np(v).astype(np.float64) for k, v in out["noise"].items()}
    ctrl = {"opt": lambda: O.OptCtrl(m64), "emp": lambda: O.EmpMeanCtrl(d, online=True), "ucb": lambda: O.UCBCtrl(d, 1.0),
            "thompson": lambda: O.ThompsonCtrl(d, std=var, sample=True, prior_mean=.5, prior_var=1 / 12.0),
            "linucb": lambda: O.LinUCBCtrl(arms, 1.0)}[kind]()
    arrays = {"reward_z": nz["reward_z"]}
    if kind == "thompson":
        arrays["thompson_z"] = nz["ctrl_z"]
    if kind == "linucb":
        arrays["linucb_first"] = nz["first_arm"][None]
    cum, meta = O.deploy_online_vec(m64, var, H, ctrl, O.ReplayNoise(arrays))
    assert np.array_equal(_np(out["context_actions"]).astype(np.float64), meta["context_actions"])
    _close(_np(out["context_rewards"]), meta["context_rewards"])
    _close(_np(out["cum_means"]), cum)
    # not materialising the context and a different shard offset change nothing
    out2 = dpt.kernels.online_loop(kind, means, H, var, seed, env_id0, materialise=False, **par)
    assert torch.equal(out["cum_means"], out2["cum_means"]) and "context_actions" not in out2
    lo = N // 3
    out3 = dpt.kernels.online_loop(kind, means[lo:], H, var, seed, env_id0 + lo, **par)
    assert torch.equal(out["cum_means"][:, lo:], out3["cum_means"])
    assert torch.equal(out["context_rewards"][lo:], out3["context_rewards"])


def test_online_loop_noise_statistics(dpt):
    means, _, _ = dpt.kernels.bandit_sample_means(20000, 5, 1, 0)
    out = dpt.kernels.online_loop("thompson", means, 50, 0.3, 1, 0, p0=0.3, p1=0.5, p2=1 / 12.0, dump=True)
    for k in ("reward_z", "ctrl_z"):
        z = out["noise"][k].double()
        n = z.numel()
        assert abs(float(z.mean())) < 5 / n ** 0.5 and abs(float(z.var()) - 1) < 6 * (2 / n) ** 0.5
    z = out["noise"]["ctrl_z"].double().reshape(-1, 5)
    c = torch.corrcoef(z.T)
    assert float((c - torch.eye(5, device=c.device)).abs().max()) < 0.01          # arms' draws are independent
    rz = out["noise"]["reward_z"].double()
    assert abs(float((rz[:-1] * rz[1:]).mean())) < 0.005                           # consecutive steps too
    out = dpt.kernels.online_loop("linucb", means, 2, 0.3, 1, 0, p0=1.0, arms=np.eye(5)[:, :2].copy(), dump=True)
    cnt = torch.bincount(out["noise"]["first_arm"].long(), minlength=5).double() / 20000
    assert float((cnt - 0.2).abs().max()) < 0.015


@pytest.mark.parametrize("kind", ["emp", "ucb"])
def test_online_loop_full_size_decisions_consistent(dpt, kind):
    """BASELINE config 4 size (10k envs x H=500, d=5): every arm the kernel chose is the argmax of the
    controller's statistic recomputed from the materialised context prefix (float64 prefix sums)."""
    N, d, H, var = 10000, 5, 500, 0.3
    means, _, _ = dpt.kernels.bandit_sample_means(N, d, 5, 0)
    out = dpt.kernels.online_loop(kind, means, H, var, 5, 0, p0=1.0, dump=True)
    a1 = out["context_actions"].double()
    z = out["noise"]["reward_z"].double().T                                                    # [N,H]
    r = (means.double()[:, None, :] * a1).sum(-1) + (0.0 + var * z)                            # float64 rewards
    _close(_np(out["context_rewards"][:, :, 0]), _np(r))
    cnt = torch.cumsum(a1, 1) - a1                                                             # before each step
    sm = torch.cumsum(a1 * r[:, :, None], 1) - a1 * r[:, :, None]
    stat = sm / cnt.clamp(min=1)
    if kind == "ucb":
        stat = stat + 1.0 / cnt.sqrt().clamp(min=1)
    pick = stat.argmax(-1)
    untried = (cnt == 0)
    first_untried = untried.double().argmax(-1)
    pick = torch.where(untried.any(-1), first_untried, pick)
    chosen = a1.argmax(-1)
    bad = pick != chosen
    top2 = stat.topk(2, dim=-1).values
    gap = (top2[..., 0] - top2[..., 1]).abs()
    assert int(bad.sum()) == int((bad & (gap < 1e-9)).sum())      # only float-order ties may differ
    assert int(bad.sum()) <= 5
    # online regret falls: late steps are better than early ones, and far better than uniform play
    reg = out["regret_sums"][:, 0] / N
    assert float(reg[-50:].mean()) < 0.5 * float(reg[:10].mean())


def test_arm_stats(dpt):
    rs = np.random.RandomState(0)
    N, Hs, d, h = 130, 50, 7, 37
    acts = rs.randint(0, d, (N, Hs))
    ca = np.eye(d, dtype=np.float32)[acts]
    cr = rs.normal(size=(N, Hs, 1)).astype(np.float32)
    sums, counts = dpt.kernels.arm_stats(torch.tensor(ca), torch.tensor(cr), h)
    want_s = np.zeros((N, d))
    want_c = np.zeros((N, d), dtype=np.int64)
    for e in range(N):
        for t in range(h):
            want_s[e, acts[e, t]] += float(cr[e, t, 0])
            want_c[e, acts[e, t]] += 1
    assert np.array_equal(_np(counts), want_c) and np.allclose(_np(sums), want_s, rtol=0, atol=1e-12)
    s0, c0 = dpt.kernels.arm_stats(torch.tensor(ca), torch.tensor(cr), 0)
    assert float(s0.abs().sum()) == 0 and int(c0.sum()) == 0
    with pytest.raises(ValueError):
        dpt.kernels.arm_stats(torch.zeros(2, 3, 40), torch.zeros(2, 3, 1))


def test_deploy_online_vec_dropin(dpt):
    """Reference call pattern (evals/eval_bandit.py:107-166) with this package's classes."""
    from dpt_b200.envs.bandit_env import BanditEnv, BanditEnvVec, LinearBanditEnv
    from dpt_b200.ctrls.ctrl_bandit import (EmpMeanPolicy, LinUCBPolicy, OptPolicy, PessMeanPolicy, ThompsonSamplingPolicy,
                                            UCBPolicy)
    from dpt_b200.evals import eval_bandit, eval_linear_bandit
    dpt.seed(0)
    rs = np.random.RandomState(0)
    N, H, d, var = 256, 60, 5, 0.3
    envs = [BanditEnv(rs.uniform(0, 1, d), H, var=var) for _ in range(N)]
    vec = BanditEnvVec(envs)
    cm = eval_bandit.deploy_online_vec(vec, OptPolicy(envs, batch_size=N), H)
    assert cm.shape == (H, N) and cm.dtype == np.float64
    assert np.allclose(cm, np.stack([e.means.max() for e in envs])[None, :], atol=1e-6)
    cm, meta = eval_bandit.deploy_online_vec(vec, UCBPolicy(envs[0], const=1.0, batch_size=N), H, include_meta=True)
    assert set(meta) == {"context_states", "context_actions", "context_next_states", "context_rewards"}
    assert meta["context_actions"].shape == (N, H, d) and meta["context_rewards"].shape == (N, H, 1)
    assert np.array_equal(meta["context_actions"][:, :d].argmax(-1), np.tile(np.arange(d), (N, 1)))   # untried arms first
    # the generic (per-step) path: same classes driven one step at a time, like the reference
    class Slow(EmpMeanPolicy):
        fused_spec = None
    np.random.seed(0)
    cm2, meta2 = eval_bandit.deploy_online_vec(vec, Slow(envs[0], online=True, batch_size=N), 12, include_meta=True)
    assert cm2.shape == (12, N) and np.array_equal(meta2["context_actions"][:, :d].argmax(-1), np.tile(np.arange(d), (N, 1)))
    # per-step decisions == recount from context (ctrls/ctrl_bandit.py:91-118) on that context
    ctrl = O.EmpMeanCtrl(d, online=True)
    for h in (5, 8, 11):
        want = ctrl.act(meta2["context_actions"][:, :h], meta2["context_rewards"][:, :h, 0], None, h)
        assert np.array_equal(want, meta2["context_actions"][:, h])
    for c in (PessMeanPolicy(envs[0], const=0.8, batch_size=N), ThompsonSamplingPolicy(envs[0], std=var, sample=False, batch_size=N)):
        c.set_batch_numpy_vec({k: v[:, :10] for k, v in meta.items()})
        a = c.act_numpy_vec(vec.reset())
        assert a.shape == (N, d) and np.all(a.sum(1) == 1)
    all_means, stats = eval_bandit.online([{"means": e.means} for e in envs], None, N, H, var)
    assert set(all_means) == {"opt", "Emp", "UCB1.0", "Thomp"} and all_means["Emp"].shape == (N, H)
    assert stats["Thomp"]["regret_mean"][-1] < stats["Emp"]["regret_mean"][-1] * 1.5
    assert stats["opt"]["regret_mean"][-1] == 0 and stats["UCB1.0"]["regret_sem"].shape == (H,)
    m, s, cm_, cs = O.regret_stats(all_means["opt"], all_means["UCB1.0"])
    assert np.allclose(stats["UCB1.0"]["regret_mean"], cm_) and np.allclose(stats["UCB1.0"]["sem"], s)
    arms = O.linear_bandit_arms(10, 2)
    trajs = [{"theta": rs.normal(0, 1, 2) / np.sqrt(2), "arms": arms} for _ in range(64)]
    am, st = eval_linear_bandit.online(trajs, None, 64, 40, var)
    assert set(am) == {"opt", "Thomp", "LinUCB"} and st["LinUCB"]["regret_mean"][-1] < 0.5 * 40 * float(np.mean(am["opt"]) - np.mean(arms @ trajs[0]["theta"]) + 1)
    from dpt_b200 import collect_data
    lenvs = [LinearBanditEnv(t["theta"], arms, 20, var=var) for t in trajs]
    cs, ca, cns, cr = collect_data.rollin_linear_bandit_vec(lenvs)
    assert cs.shape == (64, 20, 1) and ca.shape == (64, 20, 10) and cr.shape == (64, 20)
    np.random.seed(1)
    tr = collect_data.generate_linear_bandit_histories(16, 10, 2, 20, var, n_hists=1, n_samples=2, data_type="thompson")
    assert len(tr) == 32 and set(tr[0]) >= {"arms", "theta", "var", "means"} and np.array_equal(tr[0]["arms"], arms)
    lin = LinUCBPolicy(lenvs[0], const=1.0, batch_size=64)
    lin.set_batch_numpy_vec({"context_actions": ca[:, :7], "context_rewards": cr[:, :7, None]})
    want = O.LinUCBCtrl(arms, 1.0).act(ca[:, :7], cr[:, :7, None], None, 7)
    assert np.array_equal(lin.act_numpy_vec(None), want)


def test_offline_eval_matches_reference(dpt):
    """evals/eval_bandit.py offline() -- every controller sees a fixed context and plays one
    noise-free pull; rewards per env equal the reference's (Thompson sample=False consumes np.random identically)."""
    from dpt_b200.evals import eval_bandit
    from dpt_b200.models.net import Transformer
    g = golden("offline_bandit")
    d, H, var = int(g["d"]), int(g["H"]), float(g["var"])
    m = Transformer({"horizon": H, "state_dim": 1, "action_dim": d, "n_layer": int(g["n_layer"]), "n_embd": 32, "n_head": 1,
                     "dropout": 0.0, "test": True})
    m.load_state_dict({k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("sd/")}, strict=False)
    N = g["means"].shape[0]
    trajs = [{"means": g["means"][i], "context_states": np.ones((H, 1)), "context_actions": np.eye(d)[g["context_actions"][i]],
              "context_next_states": np.ones((H, 1)), "context_rewards": g["context_rewards"][i]} for i in range(N)]
    for h in (H, H // 2, 1):
        np.random.seed(int(g["seed"]) + h)
        b = eval_bandit.offline(trajs, m, N, h, var, "uniform", np_random_compat=True)
        assert set(b) == {"opt", "lnr", "emp", "thmp", "lcb"}
        for k, v in b.items():
            assert v.shape == (N,)
            assert np.allclose(v, g["h%d_%s" % (h, k)], rtol=0, atol=1e-6), (h, k)
    hs, reg = eval_bandit.offline_graph(trajs, None, N, 6, var)
    assert len(hs) == 50 and set(reg) == {"emp", "thmp", "lcb"} and all(len(v) == 50 for v in reg.values())
    assert all(np.all(v > -1e-9) for v in reg.values())


@pytest.mark.parametrize("N,d,H", [(1, 2, 1), (31, 32, 7), (33, 1, 5), (65, 3, 64), (2, 5, 129)])
def test_online_loop_edge_shapes(dpt, N, d, H):
    """Ragged sizes: one env, one step, one arm, 32 arms, N and H around the 32-wide tiles."""
    rs = np.random.RandomState(N * 100 + d)
    means = torch.tensor(rs.rand(N, d), dtype=torch.float32)
    m64 = means.double().numpy()
    for kind, par, ctrl in (("emp", dict(p0=1.0), O.EmpMeanCtrl(d, online=True)), ("ucb", dict(p0=1.0), O.UCBCtrl(d, 1.0)),
                            ("thompson", dict(p0=0.3, p1=0.5, p2=1 / 12.0), O.ThompsonCtrl(d, std=0.3, sample=True, prior_mean=.5, prior_var=1 / 12.0))):
        out = dpt.kernels.online_loop(kind, means, H, 0.3, 5, 0, dump=True, **par)
        arrays = {"reward_z": _np(out["noise"]["reward_z"]).astype(np.float64)}
        if kind == "thompson":
            arrays["thompson_z"] = _np(out["noise"]["ctrl_z"]).astype(np.float64)
        cum, meta = O.deploy_online_vec(m64, 0.3, H, ctrl, O.ReplayNoise(arrays))
        assert np.array_equal(_np(out["context_actions"]).astype(np.float64), meta["context_actions"]), kind
        _close(_np(out["context_rewards"]), meta["context_rewards"])
        reg = m64.max(1)[None] - cum
        _close(_np(out["regret_sums"])[:, 2], np.cumsum(reg, axis=0).sum(1), 1e-5)
    assert dpt.kernels.online_loop("opt", means[:0], H, 0.3, 5, 0)["cum_means"].shape == (H, 0)


@pytest.mark.parametrize("kind", ["emp", "ucb", "thompson"])
def test_regret_curves_statistical_parity(dpt, kind):
    """SURVEY §4 tier 4: regret curves from INDEPENDENT noise (Philox on the device vs numpy in the oracle)
    agree within the statistical tolerance: |mean_gpu - mean_oracle| <= 5 combined standard errors per step."""
    from dpt_b200.dist import regret_stats_from_sums
    d, H, var = 5, 40, 0.3
    N_gpu, N_or = 40000, 1500
    means, _, _ = dpt.kernels.bandit_sample_means(N_gpu, d, 21, 0)
    par = {"emp": dict(p0=1.0), "ucb": dict(p0=1.0), "thompson": dict(p0=var, p1=0.5, p2=1 / 12.0)}[kind]
    out = dpt.kernels.online_loop(kind, means, H, var, 21, 0, materialise=False, **par)
    gpu = regret_stats_from_sums(_np(out["regret_sums"]), N_gpu)
    rs = np.random.RandomState(5)
    m_or = rs.uniform(0, 1, (N_or, d))
    np.random.seed(17)
    ctrl = {"emp": O.EmpMeanCtrl(d, online=True), "ucb": O.UCBCtrl(d, 1.0),
            "thompson": O.ThompsonCtrl(d, std=var, sample=True, prior_mean=.5, prior_var=1 / 12.0)}[kind]
    cum, _ = O.deploy_online_vec(m_or, var, H, ctrl, O.GlobalNoise(record=False))
    m, s, cm, cs = O.regret_stats(np.repeat(m_or.max(1)[:, None], H, 1), cum.T)
    tol = 5 * np.sqrt(gpu["sem"] ** 2 + s ** 2) + 1e-3
    assert np.all(np.abs(gpu["mean"] - m) <= tol), np.abs(gpu["mean"] - m).max()
    tolc = 5 * np.sqrt(gpu["regret_sem"] ** 2 + cs ** 2) + 1e-3
    assert np.all(np.abs(gpu["regret_mean"] - cm) <= tolc)
    assert gpu["regret_mean"][-1] > 0.5 and gpu["mean"][-1] < gpu["mean"][d]      # it learns: late regret below early regret


@pytest.mark.parametrize("kind", ["emp", "ucb", "thompson"])
def test_online_loop_bernoulli(dpt, kind):
    """Bernoulli envs (eval.py 'bandit_bernoulli'): the fused loop draws r = [u < means[a]]; replaying the dumped
    uniforms through the oracle's controllers gives the same arms and rewards."""
    N, d, H, var, seed = 150, 5, 48, 0.3, 11
    means, _, _ = dpt.kernels.bandit_sample_means(N, d, seed, 0)
    m64 = _np(means).astype(np.float64)
    par = {"emp": dict(p0=1.0), "ucb": dict(p0=1.0), "thompson": dict(p0=var, p1=0.5, p2=1 / 12.0)}[kind]
    out = dpt.kernels.online_loop(kind, means, H, var, seed, 0, dump=True, reward_type="bernoulli", **par)
    nz = {k: _np(v).astype(np.float64) for k, v in out["noise"].items()}
    assert nz["reward_z"].min() >= 0.0 and nz["reward_z"].max() < 1.0
    ctrl = {"emp": lambda: O.EmpMeanCtrl(d, online=True), "ucb": lambda: O.UCBCtrl(d, 1.0),
            "thompson": lambda: O.ThompsonCtrl(d, std=var, sample=True, prior_mean=.5, prior_var=1 / 12.0)}[kind]()
    arrays = {"reward_z": nz["reward_z"]}
    if kind == "thompson":
        arrays["thompson_z"] = nz["ctrl_z"]
    cum, meta = O.deploy_online_vec(m64, var, H, ctrl, O.ReplayNoise(arrays), reward_type="bernoulli")
    assert np.array_equal(_np(out["context_actions"]).astype(np.float64), meta["context_actions"])
    assert np.array_equal(_np(out["context_rewards"]).astype(np.float64), meta["context_rewards"])
    _close(_np(out["cum_means"]), cum)


def test_deploy_online_vec_bernoulli_envs_fused(dpt):
    """deploy_online_vec on a BanditEnvVec of bernoulli envs takes the fused path and returns {0,1} rewards."""
    from dpt_b200.envs.bandit_env import BanditEnvVec, sample
    from dpt_b200.ctrls.ctrl_bandit import UCBPolicy
    from dpt_b200.evals.eval_bandit import deploy_online_vec
    dpt.seed(3)
    envs = [sample(5, 40, 0.3, type="bernoulli") for _ in range(64)]
    vec = BanditEnvVec(envs)
    cum, meta = deploy_online_vec(vec, UCBPolicy(envs[0], const=1.0, batch_size=64), 40, include_meta=True)
    assert cum.shape == (40, 64) and set(np.unique(meta["context_rewards"])) <= {0.0, 1.0}
    assert meta["context_actions"].sum(-1).min() == 1.0


def test_linear_bandit_offline_eval(dpt):
    """evals/eval_linear_bandit.py:202-330 `offline` / `offline_graph`: each controller sees the first h context rows
    and plays one noise-free pull.  'opt' is the best arm's mean, 'linreg' (LinUCB with const = 0) is checked against
    the oracle's LinUCB on the same context, Thompson(sample=False) against the float64 posterior (its 100-draw mode
    picks the arm with the largest posterior mean whenever the gap dwarfs the posterior std)."""
    from dpt_b200 import collect_data
    from dpt_b200.evals import eval_linear_bandit
    dpt.seed(21)
    N, d, lin_d, H, var = 24, 10, 2, 30, 0.3
    trajs = collect_data.generate_linear_bandit_histories(N, d, lin_d, H, var, n_hists=1, n_samples=1, cov=0.0, data_type="thompson")
    assert len(trajs) == N and trajs[0]["arms"].shape == (d, lin_d)
    np.random.seed(0)
    for h in (1, 7, H):
        res = eval_linear_bandit.offline(trajs, None, N, h, var)
        assert set(res) == {"opt", "thmp", "linreg"} and all(v.shape == (N,) for v in res.values())
        means = np.stack([t["means"] for t in trajs])
        _close(res["opt"], means.max(1))
        ca = np.stack([t["context_actions"][:h] for t in trajs])
        cr = np.stack([np.asarray(t["context_rewards"])[:h, None] for t in trajs])
        hot = O.LinUCBCtrl(trajs[0]["arms"], 0.0).act(ca, cr, None, h)
        _close(res["linreg"], (means * hot).sum(1))
        assert np.all(res["thmp"] <= res["opt"] + 1e-6) and np.all(res["linreg"] <= res["opt"] + 1e-6)
    hs, sub = eval_linear_bandit.offline_graph(trajs[:6], None, 6, 5, var)
    assert list(hs) == [1, 2, 3, 4, 5] and set(sub) == {"thmp", "linreg"} and sub["linreg"][0].shape == (5,) and np.all(sub["linreg"][0] >= -1e-6)


def test_deploy_online_single_env(dpt):
    """evals/eval_bandit.py:24-53 `deploy_online`: the single-env loop through set_batch / env.deploy."""
    from dpt_b200.envs.bandit_env import BanditEnv
    from dpt_b200.ctrls.ctrl_bandit import EmpMeanPolicy, OptPolicy
    from dpt_b200.evals import eval_bandit, eval_linear_bandit
    dpt.seed(2)
    env = BanditEnv(np.array([0.1, 0.8, 0.3, 0.5]), 1, var=0.1)
    cm = eval_bandit.deploy_online(env, OptPolicy(env), 6)
    assert cm.shape == (6,) and np.allclose(cm, 0.8)
    cm = eval_bandit.deploy_online(env, EmpMeanPolicy(env, online=True), 12)
    assert cm.shape == (12,) and sorted(cm[:4]) == [0.1, 0.3, 0.5, 0.8] and np.allclose(cm[4:], 0.8)   # each arm once, then the best
    assert eval_linear_bandit.deploy_online is eval_bandit.deploy_online


@pytest.fixture
def online_impl(dpt):
    """Selects the dpt_online_loop implementation for one test (dpt_debug_online_impl) and restores the default."""
    lib = dpt._lib.lib()
    yield lambda impl: lib.dpt_debug_online_impl(int(impl))
    lib.dpt_debug_online_impl(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,par,d,H,N", [("opt", {}, 5, 500, 4099), ("emp", dict(p0=1.0), 5, 500, 4099), ("ucb", dict(p0=1.0), 5, 260, 2050),
                                            ("thompson", dict(p0=0.3, p1=0.5, p2=1 / 12.0), 5, 500, 3000),
                                            ("thompson", dict(p0=0.3, p1=0.0, p2=1.0), 10, 200, 3000), ("linucb", dict(p0=1.0), 10, 200, 3000),
                                            ("emp", dict(p0=1.0), 3, 64, 777), ("ucb", dict(p0=1.0), 7, 33, 500)])
@pytest.mark.parametrize("reward_type", ["uniform", "bernoulli"])
def test_online_impls_agree(dpt, online_impl, kind, par, d, H, N, reward_type):
    """General kernel (1), fused kernel (2) and split pipeline (3) are three implementations of one contract and one Philox
    stream: identical context tensors and cum_means bit for bit; regret sums: fused vs split to 1e-9 (float64, different summation
    orders), the general kernel to 1e-6 (it stages the cumulative regret as fp32)."""
    par = dict(par)
    if kind == "linucb":
        par["arms"] = np.random.RandomState(5).normal(size=(d, 2)) / np.sqrt(2)
    means, _, _ = dpt.kernels.bandit_sample_means(N, d, 17, 0)
    outs = {}
    # (Thompson: the general kernel draws its control normals per step, the fast kernels per step pair -- different Philox
    # counters, same distribution -- so only the two fast implementations are comparable there)
    impls = (2, 3) if kind == "thompson" else (1, 2, 3)
    for impl in impls:
        online_impl(impl)
        outs[impl] = dpt.kernels.online_loop(kind, means, H, 0.3, 23, 5, reward_type=reward_type, **par)
    for impl in impls[1:]:
        for k in ("context_states", "context_actions", "context_next_states", "context_rewards", "cum_means"):
            assert torch.equal(outs[impl][k], outs[impls[0]][k]), (impl, k)
        assert torch.allclose(outs[impl]["regret_sums"], outs[impls[0]]["regret_sums"], rtol=1e-6, atol=1e-9), impl
    assert torch.allclose(outs[3]["regret_sums"], outs[2]["regret_sums"], rtol=1e-9, atol=1e-9)
    # without the context tensors (the regret sums then come from the pass over cum_means)
    online_impl(3)
    lean = dpt.kernels.online_loop(kind, means, H, 0.3, 23, 5, reward_type=reward_type, materialise=False, **par)
    assert torch.equal(lean["cum_means"], outs[2]["cum_means"])
    assert torch.allclose(lean["regret_sums"], outs[2]["regret_sums"], rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
def test_count_division_is_exact(dpt):
    """The warp-specialised loop replaces the reference's float64 ``b / max(1, counts)`` (ctrls/ctrl_bandit.py:105) by
    a table reciprocal + two FMA corrections: bit-identical quotients for every count up to 4096 and random sums."""
    from dpt_b200._lib import check, lib, ptr, stream_ptr
    rs = np.random.RandomState(0)
    n = np.tile(np.arange(1, 4097, dtype=np.int32), 64)
    a = np.concatenate([rs.normal(0, 1, n.size // 2) * n[:n.size // 2], rs.uniform(-1, 1, n.size - n.size // 2) * 10.0 ** rs.randint(-8, 8, n.size - n.size // 2)])
    a[:64] = [0.0, 1.0, -1.0, 3.0, 1e-300, 1e300, 0.1, 7.0] * 8
    ta, tn = torch.tensor(a, device="cuda"), torch.tensor(n, device="cuda")
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    check(lib().dpt_selftest_div(ptr(ta), ptr(tn), n.size, ptr(bad), stream_ptr()), "dpt_selftest_div")
    assert int(bad.item()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind,par", [("opt", {}), ("emp", dict(p0=1.0)), ("emp", dict(p0=0.0)), ("ucb", dict(p0=1.0)),
                                      ("thompson", dict(p0=0.3, p1=0.5, p2=1 / 12.0))])
@pytest.mark.parametrize("d,H,N", [(5, 500, 997), (5, 37, 65), (3, 50, 33), (10, 200, 300), (7, 19, 100), (1, 9, 40), (5, 132, 130), (4, 260, 70)])
@pytest.mark.parametrize("impl", [3])   # (fused kernel == split pipeline bit for bit: test_online_impls_agree)
def test_ws_kernel_matches_oracle_ragged(dpt, kind, par, d, H, N, impl, online_impl):
    """The fast kernels (impl 2: single fused kernel; impl 3: split pipeline = controller kernel + context expansion with the
    regret sums) against the oracle's recount-from-context controllers on the same injected noise: every arm bit for bit,
    rewards to 1e-5, cum_means exactly, the regret sums to 1e-6 -- at ragged sizes (partial warps, partial step quads,
    partial 128-step ranges, d below / at the compiled widths)."""
    online_impl(impl)
    rs = np.random.RandomState(d * 1000 + H)
    means = torch.tensor(rs.uniform(0, 1, (N, d)).astype(np.float32), device="cuda")
    inj = {"reward_z": rs.normal(0, 1, (H, N)).astype(np.float32)}
    if kind == "thompson":
        inj["ctrl_z"] = rs.normal(0, 1, (H, N, d)).astype(np.float32)
    out = dpt.kernels.online_loop(kind, means, H, 0.3, 3, 0, inject=inj, **par)
    # the oracle's recount-from-context controllers on the same noise
    mk = {"opt": lambda: O.OptCtrl(_np(means).astype(np.float64)), "emp": lambda: O.EmpMeanCtrl(d, online=par.get("p0", 0.0) != 0.0),
          "ucb": lambda: O.UCBCtrl(d, const=1.0), "thompson": lambda: O.ThompsonCtrl(d, std=0.3, sample=True, prior_mean=.5, prior_var=1 / 12.0)}[kind]
    arrays = {"reward_z": inj["reward_z"].astype(np.float64)}
    if kind == "thompson":
        arrays["thompson_z"] = inj["ctrl_z"].astype(np.float64)
    cum, meta = O.deploy_online_vec(_np(means).astype(np.float64), 0.3, H, mk(), O.ReplayNoise(arrays))
    assert np.array_equal(_np(out["context_actions"]).astype(np.float64), meta["context_actions"])
    ref_r = meta["context_rewards"][:, :, 0]
    got_r = _np(out["context_rewards"])[:, :, 0].astype(np.float64)
    assert np.all(np.abs(got_r - ref_r) <= 1e-5 * np.maximum(1.0, np.abs(ref_r)))
    assert np.array_equal(_np(out["cum_means"]).astype(np.float64), cum)
    assert bool((out["context_states"] == 1).all()) and bool((out["context_next_states"] == 1).all())
    opt = _np(means).astype(np.float64).max(1)
    reg = opt[None, :] - cum
    creg = np.cumsum(reg, 0)
    want = np.stack([reg.sum(1), (reg ** 2).sum(1), creg.sum(1), (creg ** 2).sum(1)], 1)
    assert np.allclose(_np(out["regret_sums"]), want, rtol=1e-6, atol=1e-9)
