import copy, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsc_dr_planner_b200 import capi, workloads as W
from lsc_dr_planner_b200.planner import BatchPlanner
# reproduce gpu_fuzz round 4
rng = np.random.default_rng(123)
for r in range(5):
    M, dim = [(5, 3), (5, 2), (10, 2), (10, 3)][r % 4]
    mode = int(rng.integers(0, 2)); K = int(rng.choice([0, 1, 7, 20, 40])); n = int(rng.choice([257, 1536, 2048]))
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=mode)
    batch = W.make_forest_batch(n, K=K, cfg=cfg, seed=1000 + r, moving=bool(rng.integers(0, 2)))
    if rng.random() < 0.5:
        g = batch.own_traj[:, -1, -1, :] + rng.uniform(-1.0, 1.0, (n, 3)).astype(np.float32)
        g[:, 2] = cfg.z_2d if dim == 2 else np.clip(g[:, 2], 0.3, 2.2)
        batch.goal = g.astype(np.float32)
    gen = capi.GEN_LSC if (mode == 0 or rng.random() < 0.5) else capi.GEN_CLSC
print("round", r, M, dim, mode, K, n, gen)
out = {}
for presolve in (1, 3, 0):
    c = copy.copy(batch.cfg); c.presolve = presolve
    pl = BatchPlanner(c, device=0); d = pl.upload(batch)
    pl.assemble_fused_device(d, gen); pl.solve_device(d, want_kkt=True, warm=bool(r % 3))
    torch.cuda.synchronize()
    badrow = ~torch.isfinite(d.ctrl).all(dim=1)
    idx = torch.nonzero(badrow).flatten().cpu().numpy()
    print("presolve", presolve, "non-finite agents", idx[:10], "status", d.status[badrow][:10].tolist(), "iters", d.iters[badrow][:10].tolist(), "kkt", d.kkt[badrow][:3].tolist())
    if len(idx) and not out:
        a = int(idx[0]); lo, hi = int(batch.obs_offsets[a]), int(batch.obs_offsets[a + 1])
        out = dict(agent=a, presolve=presolve, state=batch.state[a], goal=batch.goal[a], limits=batch.limits[a], own=batch.own_traj[a],
                   normals=d.normals[lo:hi].cpu().numpy(), rhs=d.rhs[lo:hi].cpu().numpy(), world=np.array(batch.cfg.world_min + batch.cfg.world_max),
                   M=M, dim=dim, mode=mode, warm=bool(r % 3), ctrl=d.ctrl[a].cpu().numpy())
os.makedirs("gpurun_out", exist_ok=True)
if out:
    np.savez("gpurun_out/nonfinite_case.npz", **out)
