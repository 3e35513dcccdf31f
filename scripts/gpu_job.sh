#!/bin/bash
mkdir -p gpurun_out
V=decision-pretrained-transformer_b200/variants
cat > /tmp/o.py <<'PY'
import sys, torch
sys.path.insert(0,'.')
import dpt_b200
from dpt_b200 import kernels
for N in (20000, 40000):
  means,_,_ = kernels.bandit_sample_means(N,5,0,0)
  for kind,par in (("opt",{}),("emp",dict(p0=1.0))):
    for mat in (True, False):
        f=lambda: kernels.online_loop(kind, means, 512, 0.3, 2, 0, materialise=mat, regret=False, **par)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/10
        print("N=%d H=512"%N, kind, "mat" if mat else "no-mat", "%.3f ms"%ms, "%.0f GB/s"%(N*512*36/ms/1e6) if mat else "", flush=True)
PY
for v in wt16 default wt64 wt128; do
  if [ $v = default ]; then unset DPT_B200_LIB; else export DPT_B200_LIB=$PWD/$V/libdpt_b200_$v.so; fi
  echo "--- $v"; python /tmp/o.py
done
