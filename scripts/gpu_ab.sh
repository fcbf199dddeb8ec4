#!/bin/bash
# GPU visit: quick parity check of the default library, then A/B of prebuilt library variants
mkdir -p gpurun_out
cp lsc_dr_planner_b200/liblscqp.so /tmp/liblscqp_default.so
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
for tag in "$@"; do
  cp lsc_dr_planner_b200/liblscqp_$tag.so lsc_dr_planner_b200/liblscqp.so
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/ab_$tag.err | tee gpurun_out/ab_$tag.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['value']), d['kernel_ms'], d['pdip_iterations_mean'], 'e2e', round(d['e2e']['value']), {k:round(v['ms'],3) for k,v in d['variants'].items() if isinstance(v,dict)})"
  tail -2 gpurun_out/ab_$tag.err
done
cp /tmp/liblscqp_default.so lsc_dr_planner_b200/liblscqp.so
