#!/bin/bash
# bench at N=1 on the GPU box; extra args go to bench.py
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "rc=$?"
python tools/show_bench.py gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_n1.err
