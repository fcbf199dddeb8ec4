"""development aid: closed loop with corridors, looks for the step where an agent gets close to an obstacle"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_missions import FOREST10
from lsc_dr_planner_b200 import missions as MS
from lsc_dr_planner_b200.closed_loop import ClosedLoopSim
from oracle import oracle as orc
boxes = np.load("tests/golden/worlds.npz")["forest1"]
mission = MS.parse_mission(FOREST10, 3, 1.0)
cfg = MS.launch_config(mission, M=5, dim=3, comm_range=0.0); cfg.use_sfc = True
batch = MS.first_replan_batch(mission, cfg); batch.goal = mission.goal.copy()
sim = ClosedLoopSim(batch, device=0, K=9, goal_mode="righthand", world_boxes=boxes)
m = orc.Map(boxes, cfg.world_min, cfg.world_max)
occ = np.argwhere(m.occupancy()); lo = (occ + np.array(m.key0)) * 0.1; hi = lo + 0.1
for step in range(80):
    sfc_before = sim.sfc.clone()
    sim.step()
    sim.sync_state()
    p = sim.state[:, :3].cpu().numpy().astype(np.float64)
    d = np.sqrt((np.maximum(np.maximum(lo[None] - p[:, None], p[:, None] - hi[None]), 0) ** 2).sum(-1))
    a = int(d.min(axis=1).argmin())
    if d.min() < 0.149:
        box = sim.sfc[a].cpu().numpy()          # corridors used by this step's QP
        traj = sim.traj_out[a].cpu().numpy()
        print("step", step, "agent", a, "dist", d.min(), "pos", p[a], "status", int(sim.status[a]), "sfc_status", int(sim.sfc_status[a]))
        print("  cube", lo[d[a].argmin()], "box0", box[0], "box1", box[1])
        for s_ in range(2):
            inside = (traj[s_] >= box[s_, :3] - 1e-5).all(axis=1) & (traj[s_] <= box[s_, 3:] + 1e-5).all(axis=1)
            print("  seg", s_, "control points inside box:", inside, traj[s_][:, :2].round(3).tolist())
        break
