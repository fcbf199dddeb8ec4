#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): bench.py under torchrun at N ranks (sharded blocks included), under a hard timeout.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_${N}.txt 2>&1
nvidia-smi topo -m > gpurun_out/topo_${N}.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; tail -5 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"]))
    for k in ("strong", "closed_loop", "closed_loop_nccl"):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e:
    print("no line:", e)
PY
