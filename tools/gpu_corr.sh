#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh tests -k "corr or smoke or streamed"
{
python tools/profile_case.py corrpairs --frames 72 --reps 20
python tools/profile_case.py corrpairs --frames 1024 --reps 10
python tools/profile_case.py corrsweep --reps 20
} > gpurun_out/corr_sweep.log 2>&1
cat gpurun_out/corr_sweep.log
