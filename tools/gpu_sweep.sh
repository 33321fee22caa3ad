#!/bin/bash
mkdir -p gpurun_out
{
for h in 512 $((512+(1<<20))) $((1<<20)); do python tools/profile_case.py fcb35 --hint $h; done
python tools/profile_case.py fcb35 --hint 512 --frames 288
python tools/profile_case.py fcb35 --hint 64 --frames 288
for c in fcb33 bb256s2 bb256 bb512s2; do for h in 64 512; do python tools/profile_case.py $c --hint $h; done; done
} > gpurun_out/sweep.log 2>&1
grep -v predictor gpurun_out/sweep.log
