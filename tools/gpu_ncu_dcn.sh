#!/bin/bash
# ncu --set full of the FCB 3x5 grouped launch for several scheduling hints
mkdir -p gpurun_out
for h in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:dcn_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_dcn_fcb35_h$h \
      python tools/profile_case.py fcb35 --hint $h --reps 2 > gpurun_out/ncu_dcn_h$h.log 2>&1
  tail -3 gpurun_out/ncu_dcn_h$h.log
done
