#!/usr/bin/env python
"""Wall time of the reference-signature drop-in calls, next to the reference's own.

    python scripts/time_dropin.py [--envs 125000]

Times ``collect_data.generate_bandit_histories(n_envs, 5, 500, 0.3, n_hists=1, n_samples=1, cov=0.0, type='uniform')``
through ``dpt_b200.install_dropin()`` (the call a user of the reference makes: returns the list of traj dicts in the
reference's dtypes) and the unmodified reference's function (baseline/_ref, one process, reduced N), prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=125000)
    ap.add_argument("--ref-envs", type=int, default=400)
    a = ap.parse_args()
    import numpy as np
    import torch
    import dpt_b200
    dpt_b200.install_dropin()
    import collect_data                     # the drop-in module under the reference's name
    H, d = 500, 5
    collect_data.generate_bandit_histories(2000, d, H, 0.3, n_hists=1, n_samples=1, cov=0.0, type="uniform")   # warm-up
    out = {"call": "collect_data.generate_bandit_histories(%d, 5, 500, 0.3, n_hists=1, n_samples=1, cov=0.0, type='uniform')" % a.envs}
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        trajs = collect_data.generate_bandit_histories(a.envs, d, H, 0.3, n_hists=1, n_samples=1, cov=0.0, type="uniform")
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        assert len(trajs) == a.envs and trajs[0]["context_actions"].dtype == np.float64
        del trajs
    out["dropin_seconds"] = best
    out["dropin_env_steps_per_s"] = a.envs * H / best
    try:
        from oracle import ref_loader
        ref = ref_loader.load()
        np.random.seed(0)
        t0 = time.perf_counter()
        ref.collect_data.generate_bandit_histories(a.ref_envs, d, H, 0.3, n_hists=1, n_samples=1, cov=0.0, type="uniform")
        dt = time.perf_counter() - t0
        out["reference_seconds_per_env"] = dt / a.ref_envs
        out["reference_env_steps_per_s_1core"] = a.ref_envs * H / dt
        out["reference_seconds_extrapolated"] = dt / a.ref_envs * a.envs
        out["speedup_vs_reference_1process"] = out["reference_seconds_extrapolated"] / best
        out["reference_sample"] = "%d envs, one process (the reference is single-threaded)" % a.ref_envs
    except Exception as e:   # noqa: BLE001
        out["reference"] = "unavailable: %s" % str(e)[:100]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
