"""CPU tier: the C-ABI library loads and exports every symbol include/dpt_b200.h declares;
argument validation / error strings work without a GPU; the package fails loudly without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dpt_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(dpt):
    lib = dpt._lib.lib()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libdpt_b200.so does not export %s" % n
        assert n in dpt._lib.PROTOTYPES, "no ctypes prototype for %s" % n
    assert sorted(dpt._lib.PROTOTYPES) == names
    assert lib.dpt_version() == dpt._lib.ABI_VERSION


def test_invalid_args_report_errors(dpt):
    lib = dpt._lib.lib()
    rc = lib.dpt_bandit_rollin(None, 0.3, 0, 0, 0, 4, 8, 99, None, None, None, None, None, None, None, None)
    assert rc == dpt._lib.ERR_INVALID_ARG and b"d=99" in lib.dpt_last_error()
    rc = lib.dpt_bandit_rollin(None, 0.3, 7, 0, 0, 4, 8, 5, None, None, None, None, None, None, None, None)
    assert rc == dpt._lib.ERR_INVALID_ARG and b"reward_type" in lib.dpt_last_error()
    rc = lib.dpt_darkroom_rollin(None, None, 10, 7, 0, 0, 4, 8, 1, None, None, None, None, None, None, None, None, None)
    assert rc == dpt._lib.ERR_INVALID_ARG and b"mode" in lib.dpt_last_error()
    with pytest.raises(ValueError):
        dpt._lib.check(rc, "dpt_darkroom_rollin")
    assert lib.dpt_bandit_rollin(None, 0.3, 0, 0, 0, 0, 8, 5, None, None, None, None, None, None, None, None) == 0  # empty batch


def test_no_cpu_fallback(dpt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(dpt._lib.DptError):
        dpt.collect_data.collect_bandit(4, 5, 8, 0.3)
    from dpt_b200.envs.bandit_env import BanditEnv, BanditEnvVec
    with pytest.raises(dpt._lib.DptError):
        BanditEnvVec([BanditEnv([0.1, 0.9], 4, var=0.1)])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "decision-pretrained-transformer_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|#include\s*[<\"].*oracle|import_module\([\"']oracle", re.M)
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dp, fn)).read()), fn


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver runs on the box's host cores) prints the contract's JSON line."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, DPT_REF_ENVS_PER_CORE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([x for x in out.stdout.splitlines() if x.startswith("{")][-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    from oracle import ref_loader
    want_kind = "reference" if ref_loader.available() else "port"    # live reference when staged (baseline/_ref), else its port
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == want_kind
    if want_kind == "reference":   # the second half of the metric rides on the same line
        assert line["online_eval"]["unit"] == "trajs/s" and line["online_eval"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["cores"] >= 1
