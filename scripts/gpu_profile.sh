#!/bin/bash
# ncu evidence for one bench command: per-launch durations + one full capture of each hot kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-sharded --no-variants"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:das_solve -s 2 -c 1 -o gpurun_out/prof_das -f $CMD > gpurun_out/ncu_das.log 2>&1
LSCQP_DAS=0 ncu --set full --clock-control none --import-source on -k regex:pdip_solve -s 4 -c 1 -o gpurun_out/prof_solve -f $CMD > gpurun_out/ncu_solve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsc_prune -s 2 -c 1 -o gpurun_out/prof_asm -f $CMD > gpurun_out/ncu_asm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lsc_pairs -s 2 -c 1 -o gpurun_out/prof_pairs -f $CMD > gpurun_out/ncu_pairs.log 2>&1
ls -la gpurun_out | head -30
