mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import dpt_b200
from dpt_b200.models.net import Transformer
from dpt_b200.envs.darkroom_env import DarkroomEnv, DarkroomEnvVec
from dpt_b200.ctrls.ctrl_darkroom import DarkroomTransformerController
from dpt_b200.evals import eval_darkroom
torch.manual_seed(0)
m = Transformer({"horizon": 100, "state_dim": 2, "action_dim": 5, "n_layer": 4, "n_embd": 32, "n_head": 1, "dropout": 0.0, "test": True})
rs = np.random.RandomState(0)
for N in (100, 1000):
    envs = [DarkroomEnv(10, rs.randint(0, 10, 2), 100) for _ in range(N)]
    vec = DarkroomEnvVec(envs)
    for prec in (0, 1):
        m.precision = prec
        eval_darkroom.deploy_online_vec_device(vec, m, 3, 100, 100)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = eval_darkroom.deploy_online_vec_device(vec, m, 40, 100, 100)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("darkroom online eval fused: N=%d Heps=40 horizon=100 H=100 precision=%d: %.3f s  (%.1f trajs/s, %.2f M env-steps/s)" % (N, prec, dt, N / dt, N * 4000 / dt / 1e6))
    if N == 100:
        m.precision = 0
        c = DarkroomTransformerController(m, batch_size=N, sample=True); c.fused = False
        t0 = time.perf_counter()
        eval_darkroom.deploy_online_vec(vec, c, 2, 100, 100)
        dt = (time.perf_counter() - t0) * 20
        print("darkroom online eval step-by-step path (extrapolated from 2 episodes): %.1f s" % dt)
PY
