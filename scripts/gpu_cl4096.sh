#!/bin/bash
# closed loop at 4096 agents (K = 40 nearest, no range filter), CUDA-graph step, N ranks
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python scripts/closed_loop_bench.py --graph --agents 4096 --steps 100 > gpurun_out/closed_loop_4096_1gpu.json 2> gpurun_out/cl4096.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/closed_loop_bench.py --graph --agents 4096 --steps 100 > gpurun_out/closed_loop_4096_${N}gpu.json 2> gpurun_out/cl4096.err
fi
tail -2 gpurun_out/cl4096.err; grep "^{" gpurun_out/closed_loop_4096_${N}gpu.json | cut -c1-500
