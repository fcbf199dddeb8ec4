#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-sharded --no-variants"
ncu --set full --clock-control none --import-source on -k regex:das_solve -s 2 -c 1 -o gpurun_out/prof_das -f $CMD > gpurun_out/ncu_das.log 2>&1
ls -la gpurun_out/prof_das.ncu-rep
