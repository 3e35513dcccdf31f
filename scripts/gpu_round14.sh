mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import dpt_b200
from dpt_b200 import kernels
from dpt_b200.models.net import Transformer
torch.manual_seed(0)
m = Transformer({"horizon": 100, "state_dim": 2, "action_dim": 5, "n_layer": 4, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
for B in (100, 4096, 32768):
    T = 100
    x = {"query_states": torch.rand(B, 2, device="cuda"), "context_states": torch.rand(B, T, 2, device="cuda"), "context_actions": torch.rand(B, T, 5, device="cuda"),
         "context_next_states": torch.rand(B, T, 2, device="cuda"), "context_rewards": torch.rand(B, T, 1, device="cuda")}
    for prec in (0, 1):
        m.precision = prec
        for _ in range(3): m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): m(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        flops = B * (101 * (98304 + 832) + 4 * 128 * 101 * 102 / 2)
        print("forward B=%d T=%d precision=%d: %.3f ms  %.1f M tokens/s  %.2f TFLOP/s" % (B, T, prec, ms, B * 101 / ms / 1e3, flops / ms / 1e9))
# online loop at a saturating size
for kind, par in (("opt", {}), ("emp", dict(p0=1.0)), ("ucb", dict(p0=1.0)), ("thompson", dict(p0=0.3, p1=0.5, p2=1/12.0))):
    N, H = 1000000, 100
    means, _, _ = kernels.bandit_sample_means(N, 5, 0, 0)
    for _ in range(2): kernels.online_loop(kind, means, H, 0.3, 1, 0, **par)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): kernels.online_loop(kind, means, H, 0.3, 1, 0, **par)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("online_loop %s N=1M H=100 d=5 materialised (incl. output allocation): %.3f ms  %.1f G env-steps/s  %.0f GB/s (%.2f of peak)" % (kind, ms, N*H/ms/1e6, N*H*36/ms/1e6, N*H*36/ms/1e6/6533.8))
PY
