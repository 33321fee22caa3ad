#!/bin/bash
# usage: gpu_ncu.sh <name> <kernel regex> <skip> -- <profile_case args...>
name=$1; regex=$2; skip=$3; shift 4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/prof_$name \
    python tools/profile_case.py "$@" > gpurun_out/ncu_$name.log 2>&1
tail -2 gpurun_out/ncu_$name.log
