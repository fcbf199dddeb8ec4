"""Per-phase device time of ClosedLoopSim.step (development aid): python scripts/closed_loop_profile.py --agents 4096"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsc_dr_planner_b200 import workloads as W
from lsc_dr_planner_b200.closed_loop import ClosedLoopSim

ap = argparse.ArgumentParser(); ap.add_argument("--agents", type=int, default=4096); ap.add_argument("--steps", type=int, default=60)
args = ap.parse_args()
batch = W.make_forest_batch(args.agents, K=40, seed=20260005, moving=False)
rng = np.random.default_rng(5)
pos = batch.state[:, :3].copy(); ang = rng.uniform(0, 2 * np.pi, args.agents)
goal = pos + np.stack([12 * np.cos(ang), 12 * np.sin(ang), np.zeros(args.agents)], 1)
half = batch.cfg.world_max[0] - 0.5; goal[:, :2] = np.clip(goal[:, :2], -half, half); batch.goal = goal.astype(np.float32)
sim = ClosedLoopSim(batch, device=0)
for _ in range(args.steps):
    sim.step()
torch.cuda.synchronize()
def timed(fn, reps=10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
print("step total       %.3f ms" % timed(sim.step))
print("neighbours       %.3f ms" % timed(sim.neighbours))
qp = sim.planner.qp; n = sim.n_local
idx = sim.neighbours(); own = sim.traj.contiguous()
print("assemble (fused) %.3f ms" % timed(lambda: qp.assemble_lsc_fused(sim.generator, sim.prune, n, own, sim.agent_meta, sim.goal, sim.state, sim.limits, sim.obs_offsets, idx, sim.traj, sim.agent_meta, sim.goal, sim.state, sim.normals, sim.rhs)))
print("solve            %.3f ms" % timed(lambda: qp.solve_batch(n, sim.state, sim.goal, sim.limits, None, sim.obs_offsets, sim.normals, sim.rhs, sim.ctrl, sim.cost, sim.status, sim.iters, initial_traj=own)))
print("iters mean %.2f  status!=0: %d" % (float(sim.iters.float().mean()), int((sim.status != 0).sum())))
print("step kernel      %.3f ms" % timed(lambda: qp.step_batch(n, sim.ctrl, sim.cfg.dt, sim.traj_out, sim.state_out, sim.shifted)))
# which pass solved the agents (0 = dual active set; otherwise its reason for deferring to the interior point), over a fresh run
import numpy as np
sim2 = ClosedLoopSim(batch, device=0, comm_range=3.0)
hist = np.zeros(8, int)
for s in range(args.steps):
    sim2.step()
    if s % 5 == 0:
        torch.cuda.synchronize()
        kl = sim2.planner.qp.last_instances(sim2.n_local)
        hist += np.bincount(kl, minlength=8)[:8]
        if (kl != 0).any() and s < 30:
            a = int(np.where(kl != 0)[0][0])
            print("  step", s, "deferred", int((kl != 0).sum()), "first agent", a, "reason", kl[a], "K", int(sim2.obs_offsets[a + 1] - sim2.obs_offsets[a]), "iters", int(sim2.iters[a]))
print("pass histogram over the run (comm_range 3.0):", hist)
