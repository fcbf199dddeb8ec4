#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for seed in 20260007; do echo "seed $seed"; python scripts/debug_das.py 4096 $seed 2>&1 | grep -E "klass hist|active rows|agent " | head -6; done
python scripts/closed_loop_profile.py --agents 1024 2>&1 | tail -8
bash scripts/gpu_ab.sh default
