#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
for m in 0 100000; do
  LSCQP_TWO_PASS_MIN=$m timeout 600 python scripts/closed_loop_bench.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('closed loop two_pass_min=$m', d['ms_per_replan_step'], d['agent_qp_per_s'], d['qp_failures_total'], d['min_safety_ratio'])"
done
for n in 512 1024 2048; do
 for m in 0 100000; do
  LSCQP_TWO_PASS_MIN=$m timeout 600 python bench.py --agents $n --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench agents=$n two_pass_min=$m', d['kernel_ms'], round(d['value']))"
 done
done
