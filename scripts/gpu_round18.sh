mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import dpt_b200
from dpt_b200.models.net import Transformer
torch.manual_seed(0)
m = Transformer({"horizon": 500, "state_dim": 1, "action_dim": 5, "n_layer": 4, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
for B in (200, 4096):
    T = 500
    x = {"query_states": torch.ones(B, 1, device="cuda"), "context_states": torch.ones(B, T, 1, device="cuda"), "context_actions": torch.rand(B, T, 5, device="cuda"),
         "context_next_states": torch.ones(B, T, 1, device="cuda"), "context_rewards": torch.rand(B, T, 1, device="cuda")}
    for _ in range(2): m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): m(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    flops = B * (501 * (98304 + 832) + 4 * 128 * 501 * 502 / 2)
    print("dense fp32 forward B=%d T=%d: %.3f ms  %.1f M tokens/s  %.2f TFLOP/s" % (B, T, ms, B * 501 / ms / 1e3, flops / ms / 1e9))
PY
