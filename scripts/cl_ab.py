import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsc_dr_planner_b200 import workloads as W
from lsc_dr_planner_b200.closed_loop import ClosedLoopSim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
def make():
    batch = W.make_forest_batch(n, K=40, seed=20260005, moving=False)
    rng = np.random.default_rng(5)
    pos = batch.state[:, :3].copy(); ang = rng.uniform(0, 2 * np.pi, n)
    goal = pos + np.stack([12 * np.cos(ang), 12 * np.sin(ang), np.zeros(n)], 1)
    half = batch.cfg.world_max[0] - 0.5; goal[:, :2] = np.clip(goal[:, :2], -half, half); batch.goal = goal.astype(np.float32)
    return ClosedLoopSim(batch, device=0)
w = make()
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.5:          # bring the clocks up before anything is timed
    w.step(); torch.cuda.synchronize()
for knn in ("", "1", "", "1"):
    for fs in ("", "1"):
        os.environ.pop("LSCQP_CL_TORCH_KNN", None); os.environ.pop("LSCQP_CL_SYNC_FAILSAFE", None)
        if knn: os.environ["LSCQP_CL_TORCH_KNN"] = "1"
        if fs: os.environ["LSCQP_CL_SYNC_FAILSAFE"] = "1"
        sim = make()
        for _ in range(3): sim.step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        its = []
        for s in range(200):
            sim.step()
            if s % 50 == 49: its.append(sim.iters.float().mean())
        b.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"torch_knn={knn or 0} sync_failsafe={fs or 0}: gpu {a.elapsed_time(b)/200:.3f} ms/step, cpu enqueue {(t1-t0)*5:.3f} ms/step, wall {(t2-t0)*5:.3f}, iters {[round(float(x),2) for x in its]}, failed {sim.failed_total}")
