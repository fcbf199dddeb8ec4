#!/bin/bash
# consolidated single-GPU visit: tests, profiles (ncu), bench (with CPU baseline), reference arm, smoke
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
bash scripts/gpu_profile.sh > /dev/null 2>&1
python scripts/ncu_summary.py $TAG > gpurun_out/ncu_summary.log 2>&1
python scripts/ncu_lines.py gpurun_out/prof_das.ncu-rep lsc_dr_planner_b200/csrc/das_kernel.cuh 30 > profiles/${TAG}_das_solve_kernel_phases.txt 2>&1      # refresh profiles/traffic.json before the bench reads it
cp profiles/traffic.json gpurun_out/traffic.json; cp profiles/${TAG}_*summary.txt profiles/${TAG}_launches.csv gpurun_out/ 2>/dev/null
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-700
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_reference.json | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
