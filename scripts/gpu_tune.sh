#!/bin/bash
for mu in 0.003 0.01 0.03 0.1 0.3 1.0; do for wd in 1e-3 1e-2; do
LSCQP_TUNE_MU0=$mu LSCQP_TUNE_WARM_DELTA=$wd python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mu0=$mu wd=$wd', d['kernel_ms']['solve'], d['pdip_iterations_mean'], 'cfg4', round(d['variants']['solve_synthetic_planes_config4']['ms'],3), d['variants']['solve_synthetic_planes_config4']['iters_mean'], d['variants']['solve_synthetic_planes_config4']['all_ok'])"
done; done
