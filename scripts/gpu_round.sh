#!/bin/bash
# consolidated single-GPU visit: tests, bench (with CPU baseline), reference arm, closed loop, profiles
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
bash scripts/gpu_profile.sh > /dev/null 2>&1
python scripts/ncu_summary.py ${1:-r01} > gpurun_out/ncu_summary.log 2>&1      # refresh profiles/traffic.json before the bench reads it
cp profiles/traffic.json gpurun_out/traffic.json
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_reference.json | cut -c1-300
timeout 600 python scripts/closed_loop_bench.py --graph > gpurun_out/closed_loop_1gpu.json 2> gpurun_out/closed_loop.err; tail -2 gpurun_out/closed_loop.err; cat gpurun_out/closed_loop_1gpu.json
timeout 600 python scripts/closed_loop_bench.py --graph --agents 4096 --steps 100 > gpurun_out/closed_loop_4096_1gpu.json 2>> gpurun_out/closed_loop.err; cat gpurun_out/closed_loop_4096_1gpu.json
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
