#!/bin/bash
# run with gpurun --gpus N: strong-scaling bench (BASELINE configs[4]) at N ranks + the 2-GPU on-device sharding tests
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" -eq 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout=600 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
echo "== N=$N rc=$?"; python tools/show_bench.py gpurun_out/scale_n$N.json | grep -v "^  "; tail -2 gpurun_out/scale_n$N.err
