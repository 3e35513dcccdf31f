"""Transformer with the reference's interface (models/net.py:9-60), inference on the CUDA path.

``Transformer(config)`` takes the reference's config dict (horizon, state_dim, action_dim, n_layer,
n_embd, n_head, dropout, test); ``forward(x)`` takes the reference's batch dict and returns
``preds[:, -1, :]`` (test) or ``preds[:, 1:, :]``.  The module's ``state_dict`` uses the reference's
key layout (HuggingFace GPT2Model under ``transformer.``, ``embed_transition.*``, ``pred_actions.*``),
so checkpoints written by the reference's train.py load unchanged.  As in the reference, ``n_head``
from the config is ignored: the trunk always has ONE head (models/net.py:29), head_dim = n_embd.

The forward pass is dpt_gpt2_forward (hand-written CUDA, fp32); this class holds no HuggingFace
dependency and no autograd path -- training is out of scope of the rollout hot path.
"""
import ctypes

import torch
import torch.nn as nn

from .. import _lib, kernels
from .._lib import check, lib, ptr, stream_ptr


class _Conv1D(nn.Module):
    """transformers.pytorch_utils.Conv1D parameter layout: weight [in, out], y = x @ W + b."""

    def __init__(self, nf, nx):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(nx, nf).normal_(std=0.02))
        self.bias = nn.Parameter(torch.zeros(nf))


class _Attn(nn.Module):
    def __init__(self, E):
        super().__init__()
        self.c_attn = _Conv1D(3 * E, E)
        self.c_proj = _Conv1D(E, E)


class _MLP(nn.Module):
    def __init__(self, E):
        super().__init__()
        self.c_fc = _Conv1D(4 * E, E)
        self.c_proj = _Conv1D(E, 4 * E)


class _Block(nn.Module):
    def __init__(self, E):
        super().__init__()
        self.ln_1 = nn.LayerNorm(E, eps=1e-5)
        self.attn = _Attn(E)
        self.ln_2 = nn.LayerNorm(E, eps=1e-5)
        self.mlp = _MLP(E)


class _GPT2Trunk(nn.Module):
    def __init__(self, n_positions, E, n_layer, vocab=50257):
        super().__init__()
        self.wte = nn.Embedding(vocab, E)       # unused by inputs_embeds, kept for checkpoint compatibility
        self.wpe = nn.Embedding(n_positions, E)
        self.h = nn.ModuleList([_Block(E) for _ in range(n_layer)])
        self.ln_f = nn.LayerNorm(E, eps=1e-5)
        nn.init.normal_(self.wte.weight, std=0.02)
        nn.init.normal_(self.wpe.weight, std=0.02)


class Transformer(nn.Module):
    """Transformer class (models/net.py:9)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.test = config["test"]
        self.horizon = config["horizon"]
        self.n_embd = config["n_embd"]
        self.n_layer = config["n_layer"]
        self.n_head = config["n_head"]
        self.state_dim = config["state_dim"]
        self.action_dim = config["action_dim"]
        self.dropout = config["dropout"]
        self.transformer = _GPT2Trunk(4 * (1 + self.horizon), self.n_embd, self.n_layer)
        self.embed_transition = nn.Linear(2 * self.state_dim + self.action_dim + 1, self.n_embd)
        self.pred_actions = nn.Linear(self.n_embd, self.action_dim)
        self._handle = None
        self._handle_version = None
        self._workspace = None
        self.precision = 0      # 0 = fp32 (reference parity 1e-5); 1 = bf16 K/V cache, fp32 arithmetic (2e-2)

    # ---- checkpoint compatibility --------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        # transformers 4.5.1 (the reference's pin) stores the causal mask buffers in the state dict
        sd = {k: v for k, v in state_dict.items() if not (k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"))}
        res = super().load_state_dict(sd, strict=strict, **kw)
        self._handle_version = None
        return res

    # ---- device handle -------------------------------------------------------------------
    def _version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def handle(self):
        """Opaque dpt_gpt2_t* for the current weights (repacked once per weight update)."""
        dev = kernels._dev()
        if next(self.parameters()).device != dev:
            self.to(dev)
        ver = self._version()
        if self._handle is not None and ver == self._handle_version:
            return self._handle
        self._free()
        t = self.transformer
        f = lambda p: p.detach().float().contiguous()   # noqa: E731
        keep = []

        def P(p):
            x = f(p)
            keep.append(x)
            return x.data_ptr()

        def arr(fn):
            a = (ctypes.c_void_p * self.n_layer)(*[P(fn(b)) for b in t.h])
            keep.append(a)
            return ctypes.cast(a, ctypes.POINTER(ctypes.c_void_p))
        w = _lib.Gpt2Weights(
            horizon=self.horizon, state_dim=self.state_dim, action_dim=self.action_dim, n_layer=self.n_layer,
            n_embd=self.n_embd, n_positions=t.wpe.weight.shape[0],
            wpe=P(t.wpe.weight), embed_w=P(self.embed_transition.weight), embed_b=P(self.embed_transition.bias),
            pred_w=P(self.pred_actions.weight), pred_b=P(self.pred_actions.bias), lnf_w=P(t.ln_f.weight), lnf_b=P(t.ln_f.bias),
            ln1_w=arr(lambda b: b.ln_1.weight), ln1_b=arr(lambda b: b.ln_1.bias),
            attn_w=arr(lambda b: b.attn.c_attn.weight), attn_b=arr(lambda b: b.attn.c_attn.bias),
            proj_w=arr(lambda b: b.attn.c_proj.weight), proj_b=arr(lambda b: b.attn.c_proj.bias),
            ln2_w=arr(lambda b: b.ln_2.weight), ln2_b=arr(lambda b: b.ln_2.bias),
            fc_w=arr(lambda b: b.mlp.c_fc.weight), fc_b=arr(lambda b: b.mlp.c_fc.bias),
            fc2_w=arr(lambda b: b.mlp.c_proj.weight), fc2_b=arr(lambda b: b.mlp.c_proj.bias))
        h = ctypes.c_void_p()
        check(lib().dpt_gpt2_create(ctypes.byref(w), ctypes.byref(h), stream_ptr()), "dpt_gpt2_create")
        torch.cuda.current_stream().synchronize()      # the D2D repack read `keep`
        self._handle, self._handle_version = h, ver
        return h

    def _free(self):
        if self._handle is not None:
            lib().dpt_gpt2_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:   # noqa: BLE001
            pass

    def _scratch(self, nbytes):
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device=kernels._dev())
        return self._workspace

    # ---- models/net.py:41-60 --------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, ctx_share=1):
        """``ctx_share`` > 1: ``query_states`` has ctx_share rows per context row (see dpt_gpt2_forward)."""
        dev = kernels._dev()
        h = self.handle()
        f = lambda t: t.to(device=dev, dtype=torch.float32)   # noqa: E731
        q = f(x["query_states"]).contiguous()
        cs, ca, cns, cr = f(x["context_states"]), f(x["context_actions"]), f(x["context_next_states"]), f(x["context_rewards"])
        B, T = ca.shape[0] * ctx_share, ca.shape[1]
        Bc = ca.shape[0]
        assert q.shape[0] == B
        # context views [:, :h] of a [B,H,.] buffer are passed with their row stride, without a copy
        stride = ca.stride(0) // max(1, ca.shape[2]) if T > 0 else 0
        ok = T > 0 and stride >= T and all(t.stride(-1) == 1 and t.stride(1) == t.shape[2] and t.stride(0) == stride * t.shape[2]
                           for t in (cs, ca, cns, cr.reshape(Bc, T, 1) if cr.dim() == 2 else cr))
        if not ok:
            cs, ca, cns, cr = cs.contiguous(), ca.contiguous(), cns.contiguous(), cr.contiguous()
            stride = T
        out = torch.empty((B, self.action_dim) if self.test else (B, T, self.action_dim), dtype=torch.float32, device=dev)
        nbytes = lib().dpt_gpt2_forward_workspace_bytes(h, B, T, self.precision)
        ws = self._scratch(nbytes)
        check(lib().dpt_gpt2_forward(h, q.data_ptr(), cs.data_ptr() if T else None, ca.data_ptr() if T else None,
                                     cns.data_ptr() if T else None, cr.data_ptr() if T else None, B, T, stride, ctx_share,
                                     1 if self.test else 0, self.precision, ptr(out), ptr(ws), ws.numel(), stream_ptr()),
              "dpt_gpt2_forward")
        return out

    # ---- step-by-step K/V-cached decode (callers that choose the pulled arm themselves) -------------
    def decoder(self, n_seqs, max_transitions=None):
        """A K/V-cached incremental view of ``forward(...)[:, -1]`` for loops that are driven from outside (the rollout
        half of train_interactive.py:97-132 / train_explorer_exploiter.py:110-166): ``dec.query(states)`` appends the
        query token (position 0) and returns the logits on an empty context; each ``dec.append(states, actions,
        next_states, rewards)`` appends one transition and returns the logits given everything appended so far --
        O(context) work per step instead of a dense forward over the whole context."""
        return _Decoder(self, n_seqs, self.horizon if max_transitions is None else max_transitions)

    # ---- fused in-context evaluation loop ---------------------------------------------------
    @torch.no_grad()
    def online_loop(self, means, horizon, var, sample, seed, env_id0=0, materialise=True, regret=True, inject=None,
                    dump=False, reward_type="uniform"):
        """deploy_online_vec with BanditTransformerController in one launch (KV-cached decode +
        sampling + env step).  Returns the same dict as kernels.online_loop."""
        dev = kernels._dev()
        h = self.handle()
        means = kernels._as(means, torch.float32, dev)
        N, d = means.shape
        assert d == self.action_dim
        H = horizon
        out = {"cum_means": torch.empty((H, N), dtype=torch.float32, device=dev)}
        if regret:
            out["regret_sums"] = torch.zeros((H, 4), dtype=torch.float64, device=dev)
        if materialise:
            out.update(context_states=torch.empty((N, H, 1), dtype=torch.float32, device=dev),
                       context_actions=torch.empty((N, H, d), dtype=torch.float32, device=dev),
                       context_next_states=torch.empty((N, H, 1), dtype=torch.float32, device=dev),
                       context_rewards=torch.empty((N, H, 1), dtype=torch.float32, device=dev))
        kv_bytes = lib().dpt_gpt2_online_kv_bytes(h, N, H, self.precision)
        kv = self._scratch(kv_bytes)
        inj_p, dump_p, keep = None, None, []
        if inject is not None:
            s = _lib.Gpt2OnlineInject()
            for k, dt in (("reward_z", torch.float32), ("ctrl_u", torch.float64)):
                if inject.get(k) is not None:
                    t = kernels._as(inject[k], dt, dev)
                    keep.append(t)
                    setattr(s, k, ptr(t))
            inj_p = ctypes.byref(s)
        noise = None
        if dump:
            noise = {"reward_z": torch.empty((H, N), dtype=torch.float32, device=dev),
                     "ctrl_u": torch.zeros((H, N), dtype=torch.float64, device=dev),
                     "logits": torch.empty((H, N, d), dtype=torch.float32, device=dev)}
            s2 = _lib.Gpt2OnlineDump()
            for k, t in noise.items():
                setattr(s2, k, ptr(t))
            dump_p = ctypes.byref(s2)
        check(lib().dpt_gpt2_online_loop(h, ptr(means), float(var), kernels.REWARD_TYPES[reward_type], 1 if sample else 0, seed, env_id0, N, H, self.precision, ptr(kv),
                                         kv.numel(), ptr(out.get("context_states")), ptr(out.get("context_actions")),
                                         ptr(out.get("context_next_states")), ptr(out.get("context_rewards")),
                                         ptr(out["cum_means"]), ptr(out.get("regret_sums")), inj_p, dump_p, stream_ptr()),
              "dpt_gpt2_online_loop")
        if noise is not None:
            out["noise"] = noise
        return out


class _Decoder:
    """See ``Transformer.decoder``.  Owns its K/V cache (per sequence [L][2][Tpad][32], fp32 or bf16 with
    ``model.precision``); valid while the query state of a sequence does not change (bandits: the constant [1])."""

    def __init__(self, model, n_seqs, max_transitions):
        self.model, self.n, self.t_max, self.pos = model, n_seqs, max_transitions + 1, 0
        self.precision = model.precision
        dev = kernels._dev()
        self._h = model.handle()
        nbytes = lib().dpt_gpt2_online_kv_bytes(self._h, n_seqs, self.t_max, self.precision)
        self.kv = torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device=dev)
        self.din = 2 * model.state_dim + model.action_dim + 1
        self._tok = torch.zeros((n_seqs, self.din), dtype=torch.float32, device=dev)

    def _step(self):
        # the handle is owned by the model and destroyed when its weights change (handle() repacks): a decoder must
        # never launch on a freed blob, and its K/V cache belongs to the old weights anyway
        if self.model.handle() is not self._h:
            raise RuntimeError("the model's weights changed since this decoder was created: build a new decoder")
        out = torch.empty((self.n, self.model.action_dim), dtype=torch.float32, device=self._tok.device)
        check(lib().dpt_gpt2_decode_step(self._h, ptr(self._tok), self.n, self.pos, self.t_max, self.precision, ptr(self.kv),
                                         self.kv.numel(), ptr(out), stream_ptr()), "dpt_gpt2_decode_step")
        self.pos += 1
        return out

    @torch.no_grad()
    def query(self, query_states):
        assert self.pos == 0, "the query token is position 0"
        dx = self.model.state_dim
        self._tok.zero_()
        self._tok[:, :dx] = torch.as_tensor(query_states, dtype=torch.float32).to(self._tok.device).reshape(self.n, dx)
        return self._step()

    @torch.no_grad()
    def append(self, states, actions, next_states, rewards):
        assert 0 < self.pos < self.t_max, "query() first; at most max_transitions appends"
        dx, du, dev = self.model.state_dim, self.model.action_dim, self._tok.device
        f = lambda a, w: torch.as_tensor(a, dtype=torch.float32).to(dev).reshape(self.n, w)   # noqa: E731
        self._tok[:, :dx] = f(states, dx)
        self._tok[:, dx:dx + du] = f(actions, du)
        self._tok[:, dx + du:2 * dx + du] = f(next_states, dx)
        self._tok[:, 2 * dx + du:] = f(rewards, 1)
        return self._step()
