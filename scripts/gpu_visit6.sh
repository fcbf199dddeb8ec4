#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for seed in 20260001 20260007; do echo "seed $seed"; python scripts/debug_das.py 4096 $seed 2>&1 | grep -E "klass hist|active rows|solve ms" | head -6; done
bash scripts/gpu_ab.sh default
