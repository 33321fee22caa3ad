#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err
echo "== N=8 rc=$?"; python tools/show_bench.py gpurun_out/scale_n8.json | grep -v "^  "; tail -2 gpurun_out/scale_n8.err
