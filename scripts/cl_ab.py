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
for graph in (False, True, False, True):
    sim = make(); sim.use_graph = graph
    for _ in range(4): sim.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for s in range(200):
        sim.step()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"n={n} graph={graph}: gpu {a.elapsed_time(b)/200:.3f} ms/step, cpu enqueue {(t1-t0)*5:.3f} ms/step, wall {(t2-t0)*5:.3f}, iters {float(sim.iters.float().mean()):.2f}, failed {sim.failed_total}, goal dist {sim.max_goal_distance():.4f}")
