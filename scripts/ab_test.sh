#!/bin/bash
# A/B two prebuilt library variants: liblscqp_<tag>.so copied over liblscqp.so in turn
for tag in "$@"; do
  cp lsc_dr_planner_b200/liblscqp_$tag.so lsc_dr_planner_b200/liblscqp.so
  python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', d['value'], d['kernel_ms'], d['pdip_iterations_mean'], {k:round(v['ms'],3) for k,v in d['variants'].items() if isinstance(v,dict)})"
done
