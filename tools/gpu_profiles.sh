#!/bin/bash
# ncu evidence for profiles/: one --set full capture per hot kernel + the launch list of a short bench run
mkdir -p gpurun_out
cap() { name=$1; regex=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/r02_$name \
      python tools/profile_case.py "$@" > gpurun_out/r02_$name.log 2>&1; tail -1 gpurun_out/r02_$name.log; }
cap dcn_fused35_f74 dcn_tc_kernel 2 fused35 --frames 74 --reps 2
cap dcn_bb256 "dcn_tc_kernel<.*0>" 2 bb256 --reps 2
cap dcn_bb128 "dcn_tc_kernel<.*0>" 2 bb128 --reps 2
cap plain_predictor_bb256 "dcn_tc_kernel<.*1>" 1 bb256 --reps 2
cap corr_pairs_960 corr_tc_kernel 2 corrpairs --frames 961 --reps 2
cap corr_sweep corr_tc_kernel 1 corrsweep --reps 2
cap roi_align roi_align_kernel 1 roialign --reps 2
# launch list of the benchmark command itself (configs[3]-sized so that ncu's serialised replay stays short)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r02_bench_under_ncu.log | cut -c1-300
