#!/bin/bash
mkdir -p gpurun_out
python scripts/debug_das.py 2>&1 | tail -20
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-sharded --no-variants"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_das.csv $CMD > gpurun_out/bench_under_ncu.log 2>&1
grep -o '"void lscqp::[a-z_]*<[^(]*([^)]*)".*' gpurun_out/launches_das.csv | awk -F'","' '{print $1, $NF}' | head -20
ncu --set full --clock-control none --import-source on -k regex:das_solve -s 2 -c 1 -o gpurun_out/prof_das -f $CMD > gpurun_out/ncu_das.log 2>&1
ls -la gpurun_out/prof_das.ncu-rep
