#!/bin/bash
# GPU visit: parity tests, then A/B of the dual active-set first pass against the interior-point passes alone
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
for tag in das nodas; do
  if [ "$tag" = "nodas" ]; then export LSCQP_DAS=0; else unset LSCQP_DAS; fi
  timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-sharded --no-variants 2>gpurun_out/ab_$tag.err | tee gpurun_out/ab_$tag.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['value']), d['kernel_ms'], d['solver_iterations_mean'], 'e2e', round(d['e2e']['value']))"
  tail -2 gpurun_out/ab_$tag.err
done
