#!/bin/bash
# GPU visit: A/B of prebuilt library variants (liblscqp_<tag>.so copied over liblscqp.so in turn); "default" = the in-tree build
mkdir -p gpurun_out
cp lsc_dr_planner_b200/liblscqp.so /tmp/liblscqp_default.so
for tag in "$@"; do
  if [ "$tag" = "default" ]; then cp /tmp/liblscqp_default.so lsc_dr_planner_b200/liblscqp.so; else cp lsc_dr_planner_b200/liblscqp_$tag.so lsc_dr_planner_b200/liblscqp.so; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-sharded --no-variants 2>gpurun_out/ab_$tag.err | tee gpurun_out/ab_$tag.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['value']), d['kernel_ms'], d['solver_iterations_mean'], 'e2e', round(d['e2e']['value']))"
  tail -2 gpurun_out/ab_$tag.err
done
cp /tmp/liblscqp_default.so lsc_dr_planner_b200/liblscqp.so
