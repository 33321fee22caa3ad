#!/bin/bash
# very last captures of round 2: the sampling kernel with everything on (chunk-major K order, patch rows, staged epilogue)
mkdir -p gpurun_out
cap() { name=$1; regex=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/r02_$name \
      python tools/profile_case.py "$@" > gpurun_out/r02_$name.log 2>&1; tail -1 gpurun_out/r02_$name.log; }
cap dcn_fused35_f74_end dcn_tc_kernel 2 fused35 --frames 74 --reps 2
cap dcn_bb128_end dcn_tc_kernel 2 bb128 --reps 2
cap dcn_bb256_end dcn_tc_kernel 2 bb256 --reps 2
