"""Timeline of one pass of the split online-loop pipeline from a -DDPT_TIMELINE build (scripts/build_variant.sh timeline
online_loop_ws.cu -DDPT_TIMELINE; DPT_B200_LIB=variants/timeline/libdpt_b200.so): start / end of every kernel, us from the first.
    python scripts/ol_timeline.py kind N H d"""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dpt_b200
from dpt_b200 import kernels, _lib
from ol_one import PAR

NAMES = {0: "fill", 30: "regret_finish"}
for c in range(8):
    NAMES[1 + c] = "ctrl[%d]" % c
    NAMES[10 + c] = "expand[%d]" % c
    NAMES[20 + c] = "regret[%d/8 of H]" % (c + 1)

if __name__ == "__main__":
    kind, N, H, d = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    par = dict(PAR[kind])
    if kind == "linucb":
        par["arms"] = torch.tensor(np.random.RandomState(1234).normal(size=(d, 2)) / np.sqrt(2), dtype=torch.float64, device="cuda")
    means, _, _ = kernels.bandit_sample_means(N, d, 0, 0)
    buf = (ctypes.c_ulonglong * 128)()
    for i in range(4):
        out = kernels.online_loop(kind, means, H, 0.3, 2, 0, **par)
        torch.cuda.synchronize()
        _lib.lib().dpt_debug_timeline(buf)
    t = np.array(buf, dtype=np.uint64).reshape(64, 2)
    live = [i for i in range(64) if t[i, 1] > 0]
    t0 = min(int(t[i, 0]) for i in live)
    print("%s N=%d H=%d d=%d chunks=%s serial=%s" % (kind, N, H, d, os.environ.get("DPT_OL_CHUNKS", "default"), os.environ.get("DPT_OL_SERIAL", "0")))
    for i in sorted(live, key=lambda i: int(t[i, 0])):
        print("  %-22s %8.1f .. %8.1f us  (%.1f)" % (NAMES.get(i, str(i)), (int(t[i, 0]) - t0) / 1e3, (int(t[i, 1]) - t0) / 1e3,
                                                    (int(t[i, 1]) - int(t[i, 0])) / 1e3))
