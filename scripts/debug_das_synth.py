"""development aid: pass histogram of the config-4 batch (synthetic random half-spaces) through the solve dispatch"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lsc_dr_planner_b200 import capi, workloads as W
from lsc_dr_planner_b200.planner import BatchPlanner
n = 4096
pop = W.make_forest_batch(n, K=40, seed=20260004)
off, nrm, rhs = W.make_synthetic_planes(pop, K=40)
dev = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
pl = BatchPlanner(pop.cfg, device=0); qp = pl.qp
state, goal, lim, own = t(pop.state), t(pop.goal), t(pop.limits), t(pop.own_traj)
offs, normals, rhs_d = t(off.astype(np.int32)), t(nrm), t(rhs)
ctrl = torch.empty((n, 90), dtype=torch.float64, device=dev); cost = torch.empty((n,), dtype=torch.float64, device=dev)
status = torch.empty((n,), dtype=torch.int32, device=dev); iters = torch.empty((n,), dtype=torch.int32, device=dev)
kkt = torch.zeros((n, 4), dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
qp.solve_batch(n, state, goal, lim, None, offs, normals, rhs_d, ctrl, cost, status, iters, kkt=kkt, stream=stream, initial_traj=own)
torch.cuda.synchronize()
kl = qp.last_instances(n); it = iters.cpu().numpy(); k2 = kkt[:, 2].cpu().numpy().astype(int)
print("klass hist", np.bincount(kl, minlength=8), "status", np.bincount(status.cpu().numpy(), minlength=5))
print("iters of active-set agents mean %.1f max %d; active rows final mean %.1f, largest max %d" % (it[kl == 0].mean() if (kl == 0).any() else 0, it[kl == 0].max() if (kl == 0).any() else 0, (k2[kl == 0] % 64).mean() if (kl == 0).any() else 0, (k2[kl == 0] // 64).max() if (kl == 0).any() else 0))
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); qp.solve_batch(n, state, goal, lim, None, offs, normals, rhs_d, ctrl, cost, status, iters, stream=stream, initial_traj=own); b.record(); torch.cuda.synchronize()
print("solve ms", a.elapsed_time(b))
