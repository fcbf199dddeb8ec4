"""development aid: status / iteration histogram of the bench batch on the GPU"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lsc_dr_planner_b200 import capi, workloads as W
from lsc_dr_planner_b200.planner import BatchPlanner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
batch = W.make_forest_batch(n, K=40)
for presolve in (1, 5, 3):
    import copy
    cfg = copy.copy(batch.cfg); cfg.presolve = presolve
    pl = BatchPlanner(cfg, device=0)
    d = pl.upload(batch)
    pl.assemble_fused_device(d)
    pl.solve_device(d, want_kkt=True)
    torch.cuda.synchronize()
    st = d.status.cpu().numpy(); it = d.iters.cpu().numpy(); kkt = d.kkt.cpu().numpy()
    try:
        kl = pl.qp.last_instances(n)
    except Exception:
        kl = np.ones(n, int)
    print("presolve", presolve, "status hist", np.bincount(st, minlength=5), "iters mean %.2f max %d" % (it.mean(), it.max()), "light", int((kl == 0).sum()))
    bad = np.where(st != 0)[0][:8]
    for a in bad:
        print("  agent", a, "status", st[a], "iters", it[a], "klass", kl[a], "kkt", kkt[a])
