"""Generate tests/golden/*.npz from the UNMODIFIED reference (dev container only).

    python -m oracle.make_golden            # from the repo root

For every case the live reference (``/root/reference`` through
``oracle/ref_loader.py``) is run under a fixed ``np.random.seed``; then the
restatement in ``oracle/dpt_oracle.py`` is run from the SAME seed through
``GlobalNoise`` and asserted bit-identical (this is the parity pin).  The noise
the run consumed is stored beside the reference outputs so the CUDA path can be
driven with identical injected noise on the GPU box, where the reference does
not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dpt_oracle as O          # noqa: E402
from oracle import ref_loader               # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _eq(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(a, b), (what, np.abs(a.astype(np.float64) - b.astype(np.float64)).max())


def bandit_rollin(ref, name, seed, n, dim, H, var):
    np.random.seed(seed)
    rt = ref.collect_data.generate_bandit_histories(n, dim, H, var, n_hists=1, n_samples=1, cov=0.0, type="uniform")
    np.random.seed(seed)
    noise = O.GlobalNoise()
    ot = O.generate_bandit_histories(n, dim, H, var, noise)
    for r, o in zip(rt, ot):
        for k in r:
            _eq(r[k], o[k], (name, k))
            assert np.asarray(r[k]).dtype == np.asarray(o[k]).dtype, (k, np.asarray(r[k]).dtype, np.asarray(o[k]).dtype)
    rec = noise.arrays()
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), dim=dim, H=H, var=var, seed=seed,
        means=rec["means"], cov_idx=rec["cov_idx"], dir_probs=rec["dir_probs"], rand_idx=rec["rand_idx"],
        u=rec["u"].reshape(n, H), z=rec["z"].reshape(n, H),
        ref_actions=np.stack([t["context_actions"] for t in rt]),
        ref_rewards=np.stack([t["context_rewards"] for t in rt]),
        ref_states=np.stack([t["context_states"] for t in rt]),
        ref_next_states=np.stack([t["context_next_states"] for t in rt]),
        ref_optimal_action=np.stack([t["optimal_action"] for t in rt]))
    print("ok", name)


def darkroom(ref, name, seed, dim, H, rollin_type, goals=None, perm_indices=None):
    np.random.seed(seed)
    if perm_indices is None:
        rt = ref.collect_data.generate_darkroom_histories(goals, dim, H, n_hists=1, n_samples=1, rollin_type=rollin_type)
    else:
        rt = ref.collect_data.generate_darkroom_permuted_histories(perm_indices, dim, H, n_hists=1, n_samples=1,
                                                                   rollin_type=rollin_type)
        goals = [[dim - 1, dim - 1]] * len(perm_indices)
    np.random.seed(seed)
    noise = O.GlobalNoise()
    ot = O.generate_mdp_histories(goals, dim, H, rollin_type, noise, perm_indices)
    for r, o in zip(rt, ot):
        assert set(r) == set(o)
        for k in r:
            _eq(r[k], o[k], (name, k))
    rec = noise.arrays()
    n = len(goals)
    kw = {}
    if rollin_type == "uniform":
        kw = dict(state=rec["state"].reshape(n, H, 2), action=rec["action"].reshape(n, H))
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), dim=dim, H=H, seed=seed, rollin_type=rollin_type,
        goals=np.asarray(goals), perm_indices=np.asarray(perm_indices if perm_indices is not None else []),
        query=rec["query"].reshape(n, 2), **kw,
        ref_states=np.stack([t["context_states"] for t in rt]),
        ref_actions=np.stack([t["context_actions"] for t in rt]).argmax(-1).astype(np.int8),
        ref_next_states=np.stack([t["context_next_states"] for t in rt]),
        ref_rewards=np.stack([t["context_rewards"] for t in rt]),
        ref_optimal_action=np.stack([t["optimal_action"] for t in rt]).argmax(-1).astype(np.int8))
    print("ok", name)


def darkroom_table(ref, name, dim):
    """Exhaustive transition / optimal-action table from the reference's own transit()."""
    goals = [(0, 0), (dim - 1, dim - 1), (3, 5), (dim - 1, 0)]
    perms = [0, 1, 57, 119]
    ns = np.zeros((len(goals), dim, dim, 5, 2), dtype=np.int64)
    rw = np.zeros((len(goals), dim, dim, 5), dtype=np.int64)
    oa = np.zeros((len(goals), dim, dim), dtype=np.int64)
    pns = np.zeros((len(perms), dim, dim, 5, 2), dtype=np.int64)
    prw = np.zeros((len(perms), dim, dim, 5), dtype=np.int64)
    poa = np.zeros((len(perms), dim, dim), dtype=np.int64)
    eye = np.eye(5)
    for g, goal in enumerate(goals):
        env = ref.darkroom_env.DarkroomEnv(dim, goal, 10)
        for x in range(dim):
            for y in range(dim):
                oa[g, x, y] = np.argmax(env.opt_action(np.array([x, y])))
                assert oa[g, x, y] == O.darkroom_opt_action_index([x, y], goal)
                for a in range(5):
                    s, r = env.transit(np.array([x, y]), eye[a])
                    ns[g, x, y, a], rw[g, x, y, a] = s, r
                    s2, r2 = O.darkroom_transit([x, y], a, goal, dim)
                    assert np.array_equal(s, s2) and r == r2
    for p, pi in enumerate(perms):
        env = ref.darkroom_env.DarkroomEnvPermuted(dim, pi, 10)
        assert tuple(env.perm) == O.DARKROOM_PERMS[pi]
        for x in range(dim):
            for y in range(dim):
                poa[p, x, y] = np.argmax(env.opt_action(np.array([x, y])))
                assert poa[p, x, y] == O.darkroom_opt_action_index([x, y], env.goal, O.DARKROOM_PERMS[pi])
                for a in range(5):
                    s, r = env.transit(np.array([x, y]), eye[a])
                    pns[p, x, y, a], prw[p, x, y, a] = s, r
    np.savez_compressed(os.path.join(OUT, name + ".npz"), dim=dim, goals=np.asarray(goals), perms=np.asarray(perms),
                        next_state=ns, reward=rw, opt_action=oa, p_next_state=pns, p_reward=prw, p_opt_action=poa,
                        perm_table=np.asarray(O.DARKROOM_PERMS))
    print("ok", name)


def _ref_online(ref, means, H, var, make_ctrl, evalmod):
    envs = [ref.bandit_env.BanditEnv(m, H, var=var) for m in means]
    vec = ref.bandit_env.BanditEnvVec(envs)
    return evalmod.deploy_online_vec(vec, make_ctrl(envs), H, include_meta=True)


def online(ref, name, seed, N, d, H, var):
    rng = np.random.RandomState(seed + 1000)
    means = rng.uniform(0, 1, (N, d)).astype(np.float32).astype(np.float64)   # fp32-exact task
    C = ref.ctrl_bandit
    cases = {
        "opt": (lambda envs: C.OptPolicy(envs, batch_size=N), lambda: O.OptCtrl(means)),
        "emp": (lambda envs: C.EmpMeanPolicy(envs[0], online=True, batch_size=N), lambda: O.EmpMeanCtrl(d, online=True)),
        "emp_offline": (lambda envs: C.EmpMeanPolicy(envs[0], online=False, batch_size=N), lambda: O.EmpMeanCtrl(d, online=False)),
        "thompson": (lambda envs: C.ThompsonSamplingPolicy(envs[0], std=var, sample=True, prior_mean=.5, prior_var=1 / 12.0,
                                                           warm_start=False, batch_size=N),
                     lambda: O.ThompsonCtrl(d, std=var, sample=True, prior_mean=.5, prior_var=1 / 12.0)),
    }
    if N == 200:   # the reference's UCB only runs at N == 200 (ctrls/ctrl_bandit.py:374)
        cases["ucb"] = (lambda envs: C.UCBPolicy(envs[0], const=1.0, batch_size=N), lambda: O.UCBCtrl(d, const=1.0))
    out = dict(means=means, H=H, var=var, seed=seed)
    for cname, (mk_ref, mk_or) in cases.items():
        np.random.seed(seed)
        rc, rmeta = _ref_online(ref, means, H, var, mk_ref, ref.eval_bandit)
        np.random.seed(seed)
        noise = O.GlobalNoise()
        oc, ometa = O.deploy_online_vec(means, var, H, mk_or(), noise)
        _eq(rc, oc, (name, cname, "cum_means"))
        for k in rmeta:
            _eq(rmeta[k], ometa[k], (name, cname, k))
        rec = noise.arrays()
        out[cname + "_reward_z"] = rec["reward_z"]                         # [H,N]
        if "thompson_z" in rec:
            out[cname + "_thompson_z"] = rec["thompson_z"]                 # [H,N,d]
        out[cname + "_actions"] = rmeta["context_actions"].argmax(-1).astype(np.int8)
        out[cname + "_rewards"] = rmeta["context_rewards"][:, :, 0]
        out[cname + "_cum_means"] = rc
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("ok", name)


def linear(ref, name, seed, N, dim, lin_d, H, var):
    arms = O.linear_bandit_arms(dim, lin_d)
    rng = np.random.RandomState(seed=1234)
    _eq(arms, rng.normal(size=(dim, lin_d)) / np.sqrt(lin_d), "arms")
    # --- Thompson collection: collect_data.py:56-80 ---
    np.random.seed(seed)
    envs = [ref.bandit_env.sample_linear(arms, H, var) for _ in range(N)]
    thetas_ref = np.stack([e.theta for e in envs])
    rs, ra, rns, rr = ref.collect_data.rollin_linear_bandit_vec(envs)
    np.random.seed(seed)
    noise = O.GlobalNoise()
    thetas = O.sample_linear_thetas(N, lin_d, noise)
    _eq(thetas_ref, thetas, "thetas")
    means = np.stack([arms @ t for t in thetas])
    _eq(np.stack([e.means for e in envs]), means, "lin means")
    oc, ometa = O.deploy_online_vec(means, var, H, O.ThompsonCtrl(dim, std=var, sample=True, prior_mean=0.0, prior_var=1.0), noise)
    _eq(ra, ometa["context_actions"], "lin thompson actions")
    _eq(rr, ometa["context_rewards"][:, :, 0], "lin thompson rewards")
    rec = noise.arrays()
    out = dict(arms=arms, thetas=thetas, means=means, H=H, var=var, seed=seed,
               thompson_reward_z=rec["reward_z"], thompson_thompson_z=rec["thompson_z"],
               thompson_actions=ra.argmax(-1).astype(np.int8), thompson_rewards=rr, thompson_cum_means=oc)
    # --- LinUCB online: evals/eval_linear_bandit.py:54-97 + ctrls/ctrl_bandit.py:491-528 ---
    np.random.seed(seed + 1)
    vec = ref.bandit_env.BanditEnvVec(envs)
    rc, rmeta = ref.eval_linear_bandit.deploy_online_vec(vec, ref.ctrl_bandit.LinUCBPolicy(envs[0], const=1.0, batch_size=N),
                                                         H, include_meta=True)
    np.random.seed(seed + 1)
    noise = O.GlobalNoise()
    oc, ometa = O.deploy_online_vec(means, var, H, O.LinUCBCtrl(arms, const=1.0), noise)
    _eq(rc, oc, "linucb cum")
    _eq(rmeta["context_actions"], ometa["context_actions"], "linucb actions")
    _eq(rmeta["context_rewards"], ometa["context_rewards"], "linucb rewards")
    rec = noise.arrays()
    out.update(linucb_reward_z=rec["reward_z"], linucb_first=rec["linucb_first"][0],
               linucb_actions=rmeta["context_actions"].argmax(-1).astype(np.int8),
               linucb_rewards=rmeta["context_rewards"][:, :, 0], linucb_cum_means=rc)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("ok", name)


def transformer(ref, name, seed, H, n_layer, n_embd, d, N, var):
    import torch
    torch.manual_seed(seed)
    cfg = {"horizon": H, "state_dim": 1, "action_dim": d, "n_layer": n_layer, "n_embd": n_embd, "n_head": 1,
           "dropout": 0.0, "test": True}
    model = ref.net.Transformer(cfg).to(ref.net.device).eval()
    with torch.no_grad():   # make LayerNorm affine / biases non-trivial so every parameter is exercised
        for k, p in model.named_parameters():
            if "wte" in k:
                continue
            if "ln_" in k and k.endswith("weight"):
                p.add_(0.2 * torch.randn_like(p))
            elif k.endswith("bias"):
                p.add_(0.1 * torch.randn_like(p))
            elif "transformer.h" in k and k.endswith("weight"):
                p.mul_(6.0)  # HF init is N(0, 0.02): scale up so attention / MLP are not ~linear
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items() if "wte" not in k and ".attn.bias" not in k
          and "masked_bias" not in k}
    rng = np.random.RandomState(seed)
    B, T = 6, H
    acts = rng.randint(0, d, (B, T))
    ctx_a = np.eye(d)[acts]
    ctx_r = rng.normal(0.5, 0.4, (B, T, 1))
    ctx_s = np.ones((B, T, 1))
    q = np.ones((B, 1))
    outs = {}
    for t in (0, 1, 5, T):
        batch = {"query_states": torch.tensor(q).float(), "zeros": torch.zeros(B, 1 + d + 1),
                 "context_states": torch.tensor(ctx_s[:, :t]).float(), "context_actions": torch.tensor(ctx_a[:, :t]).float(),
                 "context_next_states": torch.tensor(ctx_s[:, :t]).float(), "context_rewards": torch.tensor(ctx_r[:, :t]).float()}
        with torch.no_grad():
            model.test = True
            lt = model(batch).numpy()
            o = O.transformer_forward(sd, q, ctx_s[:, :t], ctx_a[:, :t], ctx_s[:, :t], ctx_r[:, :t], n_layer, test=True)
            assert np.abs(lt - o).max() < 2e-5 * max(1.0, np.abs(lt).max()), (t, np.abs(lt - o).max())
            outs["logits_t%d" % t] = lt
            if t > 0:
                model.test = False
                la = model(batch).numpy()
                o = O.transformer_forward(sd, q, ctx_s[:, :t], ctx_a[:, :t], ctx_s[:, :t], ctx_r[:, :t], n_layer, test=False)
                assert np.abs(la - o).max() < 2e-5 * max(1.0, np.abs(la).max())
                outs["logits_all_t%d" % t] = la
                model.test = True
    # --- online loop with the reference controller (sample=True): evals/eval_bandit.py:131-136 ---
    means = np.random.RandomState(seed + 7).uniform(0, 1, (N, d)).astype(np.float32).astype(np.float64)

    def logits_fn(cs, ca, cns, cr):
        batch = {"query_states": torch.ones(N, 1), "zeros": torch.zeros(N, 1 + d + 1),
                 "context_states": torch.tensor(cs).float(), "context_actions": torch.tensor(ca).float(),
                 "context_next_states": torch.tensor(cns).float(), "context_rewards": torch.tensor(cr).float()}
        with torch.no_grad():
            return model(batch).numpy()
    np.random.seed(seed)
    rc, rmeta = _ref_online(ref, means, H, var, lambda envs: ref.ctrl_bandit.BanditTransformerController(model, sample=True, batch_size=N),
                            ref.eval_bandit)
    np.random.seed(seed)
    noise = O.GlobalNoise()
    oc, ometa = O.deploy_online_vec(means, var, H, O.TransformerCtrl(logits_fn, d, sample=True), noise)
    _eq(rc, oc, "transformer cum")
    _eq(rmeta["context_actions"], ometa["context_actions"], "transformer actions")
    _eq(rmeta["context_rewards"], ometa["context_rewards"], "transformer rewards")
    rec = noise.arrays()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), H=H, n_layer=n_layer, n_embd=n_embd, d=d, var=var, seed=seed,
                        fwd_actions=acts.astype(np.int8), fwd_rewards=ctx_r, online_means=means,
                        online_ctrl_u=rec["ctrl_u"].reshape(H, N), online_reward_z=rec["reward_z"],
                        online_actions=rmeta["context_actions"].argmax(-1).astype(np.int8),
                        online_rewards=rmeta["context_rewards"][:, :, 0], online_cum_means=rc,
                        **{"sd/" + k: v for k, v in sd.items()}, **outs)
    print("ok", name)


def darkroom_online(ref, name, seed, dim, horizon, H, Heps, n_layer, goals=None, perm_indices=None):
    import torch
    torch.manual_seed(seed)
    cfg = {"horizon": H, "state_dim": 2, "action_dim": 5, "n_layer": n_layer, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True}
    model = ref.net.Transformer(cfg).to(ref.net.device).eval()
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "wte" in k:
                continue
            if "ln_" in k and k.endswith("weight"):
                p.add_(0.2 * torch.randn_like(p))
            elif k.endswith("bias"):
                p.add_(0.1 * torch.randn_like(p))
            elif "transformer.h" in k and k.endswith("weight"):
                p.mul_(5.0)
            elif "embed_transition.weight" in k:
                p.mul_(0.3)       # states go up to dim-1: keep activations O(1)
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items() if "wte" not in k and ".attn.bias" not in k
          and "masked_bias" not in k}
    if perm_indices is None:
        envs = [ref.darkroom_env.DarkroomEnv(dim, g, horizon) for g in goals]
    else:
        envs = [ref.darkroom_env.DarkroomEnvPermuted(dim, int(pi), horizon) for pi in perm_indices]
        goals = [[dim - 1, dim - 1]] * len(perm_indices)
    N = len(envs)
    np.random.seed(seed)
    ctrl = ref.ctrl_darkroom.DarkroomTransformerController(model, batch_size=N, sample=True)
    rc = ref.eval_darkroom.deploy_online_vec(ref.darkroom_env.DarkroomEnvVec(envs), ctrl, Heps, H, horizon)

    def logits_fn(q, cs, ca, cns, cr):
        batch = {"query_states": torch.tensor(q).float(), "zeros": torch.zeros(N, 10),
                 "context_states": torch.tensor(cs).float(), "context_actions": torch.tensor(ca).float(),
                 "context_next_states": torch.tensor(cns).float(), "context_rewards": torch.tensor(cr).float()}
        with torch.no_grad():
            lt = model(batch).numpy()
        o = O.transformer_forward(sd, q, cs, ca, cns, cr, n_layer, test=True)
        assert np.abs(lt - o).max() < 3e-5 * max(1.0, np.abs(lt).max()), np.abs(lt - o).max()
        return lt
    np.random.seed(seed)
    noise = O.GlobalNoise()
    oc, _ = O.deploy_online_vec_darkroom(goals, dim, Heps, H, horizon, logits_fn, noise, perm_indices)
    _eq(rc, oc, (name, "returns"))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), dim=dim, horizon=horizon, H=H, Heps=Heps, n_layer=n_layer, seed=seed,
                        goals=np.asarray(goals), perm_indices=np.asarray(perm_indices if perm_indices is not None else []),
                        ctrl_u=noise.arrays()["ctrl_u"].reshape(Heps * horizon, N), ref_returns=rc,
                        **{"sd/" + k: v for k, v in sd.items()})
    print("ok", name)


def offline_eval(ref, name, seed, N, d, H, var, n_layer):
    """evals/eval_bandit.py:214-301 (offline) on trajectories from the reference's own collection."""
    import torch
    torch.manual_seed(seed)
    cfg = {"horizon": H, "state_dim": 1, "action_dim": d, "n_layer": n_layer, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True}
    model = ref.net.Transformer(cfg).to(ref.net.device).eval()
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "wte" not in k:
                p.add_(0.15 * torch.randn_like(p))
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items() if "wte" not in k and ".attn.bias" not in k
          and "masked_bias" not in k}
    np.random.seed(seed)
    trajs = ref.collect_data.generate_bandit_histories(N, d, H, var, n_hists=1, n_samples=1, cov=0.0, type="uniform")
    for t in trajs:   # fp32-exact tasks so both sides see identical means
        t["means"] = t["means"].astype(np.float32).astype(np.float64)
    out = {}
    for h in (H, H // 2, 1):
        np.random.seed(seed + h)
        b = ref.eval_bandit.offline(trajs, model, N, h, var, "uniform")
        for k, v in b.items():
            out["h%d_%s" % (h, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), d=d, H=H, var=var, seed=seed, n_layer=n_layer,
                        means=np.stack([t["means"] for t in trajs]),
                        context_actions=np.stack([t["context_actions"] for t in trajs]).argmax(-1).astype(np.int8),
                        context_rewards=np.stack([t["context_rewards"] for t in trajs]),
                        **out, **{"sd/" + k: v for k, v in sd.items()})
    print("ok", name)


def gpu_bandit_env(ref, name, seed, N, d, H, var):
    """envs/gpu_bandit_env.py:12-82 run on CPU with torch.randn / torch.bernoulli patched to record the noise they
    return (bernoulli: the uniforms behind torch's own rule ``u < p``), H steps + the ValueError of step H+1."""
    import torch
    out = dict(N=N, d=d, H=H, var=var, seed=seed)
    G = ref.gpu_bandit_env
    for typ in ("uniform", "bernoulli"):
        torch.manual_seed(seed)
        env = G.GPUBanditEnv(d, N, H, var=var, type=typ, device=torch.device("cpu"))
        attrs = {k: getattr(env, k) for k in ("dims", "dim", "n_envs", "H_context", "H", "var", "dx", "du", "topk", "type")}
        assert attrs == dict(dims=d, dim=d, n_envs=N, H_context=H, H=H, var=var, dx=1, du=d, topk=False, type=typ)
        out[typ + "_means"] = env.means.numpy().copy()
        out[typ + "_opt_a_index"] = env.opt_a_index.numpy().copy()
        out[typ + "_opt_a"] = env.opt_a.numpy().copy()
        rec = []
        real_randn, real_bern = torch.randn, torch.bernoulli

        def randn(*a, **k):
            z = real_randn(*a, **k)
            rec.append(z.numpy().copy())
            return z

        def bernoulli(p_, *a, **k):
            u = torch.rand(p_.shape)
            rec.append(u.numpy().copy())
            return (u < p_).to(p_.dtype)
        torch.randn, torch.bernoulli = randn, bernoulli
        try:
            s0 = env.reset()
            assert s0.shape == (N, 1) and bool((s0 == 1).all())
            acts, rs, dones = [], [], []
            g = torch.Generator().manual_seed(seed + 1)
            for t in range(H):
                a = torch.nn.functional.one_hot(torch.randint(0, d, (N,), generator=g), d).float()
                st, r, done, info = env.step(a)
                assert st.shape == (N, 1) and bool((st == 1).all()) and info == {} and done.dtype == torch.bool
                acts.append(a.numpy().argmax(1)), rs.append(r.numpy().copy()), dones.append(done.numpy().copy())
            try:
                env.step(a)
                raise AssertionError("reference did not raise past H")
            except ValueError as e:
                out[typ + "_error"] = str(e)
            out[typ + "_arm_value"] = env.get_arm_value(a).numpy().copy()
            out[typ + "_last_action"] = a.numpy().argmax(1)
        finally:
            torch.randn, torch.bernoulli = real_randn, real_bern
        out[typ + "_actions"] = np.stack(acts).astype(np.int8)
        out[typ + "_noise"] = np.stack(rec)
        out[typ + "_rewards"] = np.stack(rs)
        out[typ + "_done"] = np.stack(dones)
        # the numpy restatement of transit (:53-63) reproduces the recorded rewards from the recorded noise
        m = out[typ + "_means"]
        for t in range(H):
            ma = m[np.arange(N), acts[t]]
            want = (ma + rec[t] * np.float32(var)) if typ == "uniform" else (rec[t] < ma).astype(np.float32)
            _eq(rs[t], want.astype(np.float32), (name, typ, t))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("ok", name)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    bandit_rollin(ref, "bandit_rollin_d5", seed=0, n=8, dim=5, H=24, var=0.3)
    bandit_rollin(ref, "bandit_rollin_d10", seed=3, n=5, dim=10, H=37, var=0.1)
    bandit_rollin(ref, "bandit_rollin_d3", seed=5, n=4, dim=3, H=16, var=0.0)
    goals = [(0, 0), (9, 9), (3, 5), (9, 0), (4, 4), (0, 7)]
    darkroom(ref, "darkroom_uniform", seed=0, dim=10, H=40, rollin_type="uniform", goals=goals)
    darkroom(ref, "darkroom_expert", seed=1, dim=10, H=25, rollin_type="expert", goals=goals)
    darkroom(ref, "darkroom_perm_uniform", seed=2, dim=7, H=30, rollin_type="uniform", perm_indices=[0, 1, 57, 119, 33])
    darkroom(ref, "darkroom_perm_expert", seed=3, dim=7, H=20, rollin_type="expert", perm_indices=[0, 1, 57, 119, 33])
    darkroom_table(ref, "darkroom_table", dim=10)
    online(ref, "online_d5_n200", seed=0, N=200, d=5, H=30, var=0.3)
    online(ref, "online_d10_n16", seed=2, N=16, d=10, H=40, var=0.1)
    linear(ref, "linear_bandit", seed=4, N=24, dim=10, lin_d=2, H=30, var=0.3)
    transformer(ref, "transformer_l2", seed=0, H=12, n_layer=2, n_embd=32, d=5, N=8, var=0.3)
    transformer(ref, "transformer_l4", seed=1, H=40, n_layer=4, n_embd=32, d=5, N=4, var=0.3)
    darkroom_online(ref, "darkroom_online", seed=2, dim=6, horizon=8, H=16, Heps=5, n_layer=2,
                    goals=[(0, 0), (5, 5), (2, 3), (5, 0), (1, 4), (3, 3), (0, 5)])
    darkroom_online(ref, "darkroom_online_perm", seed=3, dim=5, horizon=6, H=12, Heps=4, n_layer=3, perm_indices=[0, 7, 57, 119])
    offline_eval(ref, "offline_bandit", seed=6, N=40, d=5, H=20, var=0.3, n_layer=2)
    gpu_bandit_env(ref, "gpu_bandit_env", seed=8, N=64, d=5, H=4, var=0.3)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "offline":
        offline_eval(ref_loader.load(), "offline_bandit", seed=6, N=40, d=5, H=20, var=0.3, n_layer=2)
    elif len(sys.argv) > 1 and sys.argv[1] == "gpu_bandit_env":
        gpu_bandit_env(ref_loader.load(), "gpu_bandit_env", seed=8, N=64, d=5, H=4, var=0.3)
    elif len(sys.argv) > 1 and sys.argv[1] == "darkroom_online":
        ref_ = ref_loader.load()
        darkroom_online(ref_, "darkroom_online", seed=2, dim=6, horizon=8, H=16, Heps=5, n_layer=2,
                        goals=[(0, 0), (5, 5), (2, 3), (5, 0), (1, 4), (3, 3), (0, 5)])
        darkroom_online(ref_, "darkroom_online_perm", seed=3, dim=5, horizon=6, H=12, Heps=4, n_layer=3, perm_indices=[0, 7, 57, 119])
    else:
        main()
