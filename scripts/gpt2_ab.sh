#!/usr/bin/env bash
# A/B the GPT-2 online-loop kernels across builds in variants/*.so (same ABI, different launch bounds)
shopt -s nullglob
for lib in "" variants/*.so; do
  echo "== ${lib:-default}"
  DPT_B200_LIB=${lib:+$PWD/$lib} timeout -s KILL 200 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, dpt_b200
from dpt_b200 import kernels
from dpt_b200.models.net import Transformer
torch.manual_seed(0)
for prec in (0, 1):
    m = Transformer({"horizon": 500, "state_dim": 1, "action_dim": 5, "n_layer": 4, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
    m.precision = prec
    means, _, _ = kernels.bandit_sample_means(10000, 5, 0, 0)
    m.online_loop(means, 500, 0.3, True, 0, 0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(3):
        m.online_loop(means, 500, 0.3, True, 1 + i, 0)
    b.record(); torch.cuda.synchronize()
    print("precision", prec, "ms", round(a.elapsed_time(b) / 3, 2))
PY
done
