#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --no-e2e --no-cpu-baseline "$@" > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
cat gpurun_out/bench_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'])
for k in ('roofline','roofline_correlation'):
    r=d.get(k)
    if r: print(k, r['kernel'], 'ms', r['ms_per_launch'], 'ach', r['achieved'], 'frac', r['frac'])
print(d.get('clocks'))
"
tail -5 gpurun_out/bench_quick.err
