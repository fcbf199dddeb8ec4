"""development aid: a small replan through both active-set instances, to be run under compute-sanitizer
(memcheck / racecheck): python scripts/sanitize_das.py N   (N <= 1332: the large instance alone; above: both)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lsc_dr_planner_b200 import workloads as W
from lsc_dr_planner_b200.planner import BatchPlanner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch = W.make_forest_batch(n, K=40)
pl = BatchPlanner(batch.cfg, device=0)
d = pl.upload(batch)
pl.assemble_fused_device(d)
pl.solve_device(d, want_kkt=True)
torch.cuda.synchronize()
kl = pl.qp.last_instances(n)
print("n", n, "status", np.bincount(d.status.cpu().numpy(), minlength=5), "first-pass share", float((kl == 0).mean()))
