#!/bin/bash
# GPU visit: A/B of prebuilt library variants (liblscqp_<tag>.so copied over liblscqp.so in turn)
mkdir -p gpurun_out
cp lsc_dr_planner_b200/liblscqp.so /tmp/liblscqp_default.so
for tag in "$@"; do
  cp lsc_dr_planner_b200/liblscqp_$tag.so lsc_dr_planner_b200/liblscqp.so
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/ab_$tag.err | tee gpurun_out/ab_$tag.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['value']), d['kernel_ms'], d['pdip_iterations_mean'], 'e2e', round(d['e2e']['value']), {k:round(v['ms'],3) for k,v in d['variants'].items() if isinstance(v,dict)})"
  tail -2 gpurun_out/ab_$tag.err
done
cp /tmp/liblscqp_default.so lsc_dr_planner_b200/liblscqp.so
