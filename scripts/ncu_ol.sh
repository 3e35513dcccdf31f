#!/bin/bash
# ncu --set full of the split online-loop kernels (controller + expander) of one configuration; the report comes back in gpurun_out/ (summarise with scripts/ncu_summary.py)
#   scripts/ncu_ol.sh <tag> <kind> <N> <H> <d>
tag=$1; shift
ncu --set full --clock-control none --import-source on -k regex:"online_ctrl_kernel|online_expand_kernel|online_loop_ws_kernel" --launch-skip ${NCU_SKIP:-4} --launch-count ${NCU_COUNT:-2} \
  -o gpurun_out/$tag -f python scripts/ol_one.py "$@" 4 > gpurun_out/$tag.log 2>&1
