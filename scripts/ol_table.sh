#!/bin/bash
# online-loop kernel table, materialised rows only, compact: name ms frac
python scripts/bench_kernels.py --only online --reps ${REPS:-12} 2>&1 | grep "materialise=True" | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('  ', j['kernel'].replace('online_loop ','').replace(' materialise=True',''), round(j['ms_mean'],4), round(j['frac_of_measured_hbm_peak'],3))
"
