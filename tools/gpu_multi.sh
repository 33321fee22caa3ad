#!/bin/bash
# run with gpurun --gpus N: multi-GPU test + scaling bench at N (c5, strong scaling)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout=600 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1
tail -5 gpurun_out/pytest_multi.log
for n in $(seq 1 $N); do
  if [ $n -eq 1 ] || [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; then
    if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 10 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --steps 10 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
    echo "== N=$n rc=$?"; python tools/show_bench.py gpurun_out/scale_n$n.json | grep -v "^  "; tail -2 gpurun_out/scale_n$n.err
  fi
done
