#!/bin/bash
# round-2 (second half) ncu evidence: the TMA shifted-view convolution (predictor + head conv), the FCB launch in its chunk-major
# K order, and the launch list of a short bench run
mkdir -p gpurun_out
cap() { name=$1; regex=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/r02_$name \
      python tools/profile_case.py "$@" > gpurun_out/r02_$name.log 2>&1; tail -1 gpurun_out/r02_$name.log; }
cap tma_predictor_bb256 conv_tma_kernel 1 bb256 --frames 1024 --reps 2
cap tma_predictor_bb128s2 conv_tma_kernel 1 bb128s2 --frames 256 --reps 2
cap tma_head_conv conv_tma_kernel 1 headconv --frames 128 --reps 2
cap dcn_fused35_f74_final dcn_tc_kernel 2 fused35 --frames 74 --reps 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_bench_launches.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r02b_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r02b_bench_under_ncu.log | cut -c1-300
