#!/bin/bash
# end-of-round captures of the backbone sampling launches (plane-major offsets, chunk-major K order, 8x8 patch rows)
mkdir -p gpurun_out
cap() { name=$1; regex=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/r02_$name \
      python tools/profile_case.py "$@" > gpurun_out/r02_$name.log 2>&1; tail -1 gpurun_out/r02_$name.log; }
cap dcn_bb128_final dcn_tc_kernel 2 bb128 --reps 2
cap dcn_bb256_final dcn_tc_kernel 2 bb256 --reps 2
