#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/closed_loop_profile.py --agents 1024 2>&1 | tail -12
timeout 900 python bench.py --no-cpu --no-variants > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", d["ms_per_step"], d["kernel_ms"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
for k in ("strong", "closed_loop"):
    print(k, {kk: d[k].get(kk) for kk in ("ms_per_step", "ms_per_replan", "value", "agent_qp_per_s", "qp_failures", "all_solved", "min_safety_ratio_end")})
PY
