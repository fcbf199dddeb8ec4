"""Randomised sweep, part 2 (development aid): communication-range (dense / compact instances), SFC boxes, goal LP."""
import copy, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsc_dr_planner_b200 import capi, workloads as W
from lsc_dr_planner_b200.planner import BatchPlanner

rng = np.random.default_rng(7)
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    M, dim = [(10, 2), (5, 3), (5, 2), (10, 3)][r % 4]
    comm = float(rng.choice([0.7, 1.5, 3.0]))
    K = int(rng.choice([3, 9]))
    n = int(rng.choice([130, 700]))
    use_sfc = bool(r % 2)
    base = W.PlannerConfig(M=M, dim=dim, planner_mode=1, comm_range=comm, use_sfc=use_sfc)
    batch = W.make_forest_batch(n, K=K, cfg=base, seed=2000 + r, moving=bool(rng.integers(0, 2)))
    last = batch.own_traj[:, -1, -1, :]
    batch.goal = (last + rng.uniform(-0.5, 0.5, last.shape)).astype(np.float32)
    batch.next_waypoint = (last + rng.uniform(-0.8, 0.8, last.shape)).astype(np.float32)
    if dim == 2:
        batch.goal[:, 2] = base.z_2d; batch.next_waypoint[:, 2] = base.z_2d
    if use_sfc:
        lo = batch.own_traj.min(axis=2) - rng.uniform(0.1, 0.8, (n, M, 3)); hi = batch.own_traj.max(axis=2) + rng.uniform(0.1, 0.8, (n, M, 3))
        batch.sfc = np.ascontiguousarray(np.concatenate([lo, hi], axis=2).astype(np.float32))
    res = []
    for max_obs in (9, 40):
        c = copy.copy(batch.cfg); c.max_obs = max_obs
        pl = BatchPlanner(c, device=0); d = pl.upload(batch)
        pl.plan_device(d, capi.GEN_CLSC)
        torch.cuda.synchronize()
        res.append((d.ctrl.clone(), d.status.clone(), d.goal_status.clone(), d.goal.clone(), float(d.iters.float().mean())))
    ok = (res[0][1] == 0) & (res[1][1] == 0)
    same_status = bool(torch.equal(res[0][1] == 0, res[1][1] == 0))
    dd = float((res[0][0][ok] - res[1][0][ok]).abs().max()) if ok.any() else 0.0
    fin = bool(torch.isfinite(res[0][0]).all() and torch.isfinite(res[1][0]).all() and torch.isfinite(res[0][3]).all())
    print(f"round {r:2d} M{M} D{dim} comm{comm} K{K} n{n} sfc{int(use_sfc)}: solved {int(ok.sum())}/{n}, goal-LP infeasible {int((res[0][2] != 0).sum())}, "
          f"same status {same_status}, compact-vs-full max diff {dd:.2e}, iters {res[0][4]:.2f}/{res[1][4]:.2f}, finite {fin}")
