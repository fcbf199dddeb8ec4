"""Randomised robustness sweep on the GPU (development aid): many workload shapes through three dispatches -- the default
(dual active-set passes first), the interior-point instances alone, and those without presolve; reports failures,
self-certified KKT residuals, the share of agents the active-set passes solved and the agreement between the dispatches.  python scripts/gpu_fuzz.py [n_rounds]"""
import copy, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from lsc_dr_planner_b200 import capi, workloads as W
from lsc_dr_planner_b200.planner import BatchPlanner

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 24
TOL = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
rng = np.random.default_rng(123)
tot = bad = 0
worst = dict(prim=0.0, gap=0.0, stat=0.0, diff=0.0)
for r in range(rounds):
    M, dim = [(5, 3), (5, 2), (10, 2), (10, 3)][r % 4]
    mode = int(rng.integers(0, 2))
    K = int(rng.choice([0, 1, 7, 20, 40]))
    n = int(rng.choice([257, 1536, 2048]))
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=mode, tol=TOL)
    batch = W.make_forest_batch(n, K=K, cfg=cfg, seed=1000 + r, moving=bool(rng.integers(0, 2)))
    if rng.random() < 0.5:
        g = batch.own_traj[:, -1, -1, :] + rng.uniform(-1.0, 1.0, (n, 3)).astype(np.float32)
        g[:, 2] = cfg.z_2d if dim == 2 else np.clip(g[:, 2], 0.3, 2.2)
        batch.goal = g.astype(np.float32)
    gen = capi.GEN_LSC if (mode == 0 or rng.random() < 0.5) else capi.GEN_CLSC
    res = []
    share = 0.0
    for presolve in (1, 9, 8):
        c = copy.copy(batch.cfg); c.presolve = presolve
        pl = BatchPlanner(c, device=0)
        d = pl.upload(batch)
        pl.assemble_fused_device(d, gen)
        pl.solve_device(d, want_kkt=True, warm=bool(r % 3))
        torch.cuda.synchronize()
        if presolve == 1:
            try:
                share = float((pl.qp.last_instances(n) == 0).mean())
            except Exception:
                share = -1.0
        res.append((d.ctrl.clone(), d.status.clone(), d.kkt.clone(), float(d.iters.float().mean())))
    s = [x[1] for x in res]
    ok = (s[0] == 0) & (s[1] == 0) & (s[2] == 0)
    nbad = int((~ok).sum()); bad += nbad; tot += n
    k = res[0][2][ok]
    dd = max(float((res[0][0][ok] - res[1][0][ok]).abs().max()) if ok.any() else 0.0,
             float((res[0][0][ok] - res[2][0][ok]).abs().max()) if ok.any() else 0.0)
    worst["prim"] = max(worst["prim"], float(k[:, 1].max())); worst["gap"] = max(worst["gap"], float(k[:, 3].max()))
    worst["stat"] = max(worst["stat"], float(k[:, 0].max())); worst["diff"] = max(worst["diff"], dd)
    fin = all(bool(torch.isfinite(x[0]).all()) for x in res)
    dfull = (res[0][0] - res[1][0]).abs().amax(dim=1); dnop = (res[0][0] - res[2][0]).abs().amax(dim=1)
    big = ok & ((dfull > 1e-5) | (dnop > 1e-5))
    if int(big.sum()):
        i = int(torch.nonzero(big)[0])
        print(f"    {int(big.sum())} agents differ by > 1e-5 (vs full: {int((ok & (dfull > 1e-5)).sum())}, vs no-presolve: {int((ok & (dnop > 1e-5)).sum())}); "
              f"agent {i}: kkt two-pass {res[0][2][i].tolist()} no-presolve {res[2][2][i].tolist()}")
    print(f"round {r:2d} M{M} D{dim} mode{mode} gen{gen} K{K:2d} n{n}: active-set share {share:.3f}, not-ok {nbad} (statuses default/interior-point/no-presolve "
          f"{[int((x != 0).sum()) for x in s]}), iters {res[0][3]:.2f}/{res[1][3]:.2f}/{res[2][3]:.2f}, max diff {dd:.2e}, finite {fin}")
print("total agents", tot, "not ok", bad, "worst", worst)
