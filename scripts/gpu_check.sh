#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -25
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", d["ms_per_step"], d["kernel_ms"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], "clocks", d["clocks"])
for k in ("strong", "closed_loop", "cpu_baseline"):
    print(k, json.dumps(d.get(k))[:900])
print({k: (round(v["ms"], 3) if isinstance(v, dict) and "ms" in v else v) for k, v in d["variants"].items()})
PY
