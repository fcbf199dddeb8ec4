#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1]); print(round(d['value']), d['kernel_ms'], 'e2e', round(d['e2e']['value']), {k: round(v['ms'],3) for k,v in d['variants'].items() if isinstance(v, dict)})"
python scripts/closed_loop_profile.py --agents 4096 2>&1 | tail -8
python scripts/cl_ab.py 1024 2>&1 | tail -2
