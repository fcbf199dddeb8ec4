#!/usr/bin/env python
"""Fly one of the reference's mission files in the batched closed loop and write the reference's result files
(log/simulation_*.csv rows and the log/summary_*.csv row).  The map pipeline and the grid MAPF layer are outside this
repository's scope: agents head for their desired goals under LSC constraints (no static obstacles, no SFC) with one of
the reference's grid-free goal modes: "righthand" (goalPlanningWithRightHandRule, src/traj_planner.cpp:468-477 with
isDeadlock :904-923) or "static".  Measured on forest10 (10 agents swapping across a circle of radius 4):
  righthand  56 replans = 11.2 s flight time, 85.5 m total distance, agent safety ratio 1.00002, no QP failure, every
             agent within 0.1 m of its goal  (the reference's own run of this mission, with static obstacles, SFC and
             grid-based goals: 15.8 s, 103.2 m, 1.021 -- log/summary_LSC_10agents.csv)
  static     deadlocks around the centre, as the symmetric swap must without a deadlock rule (600 replans, safety
             ratio 1.000003, no QP failure, agents 4.6 m short of their goals)

  python scripts/run_mission.py missions/forest10/forest10_1.json --out gpurun_out/mission
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

FOREST10 = {"quadrotors": {"crazyflie": {"max_vel": [1.0, 1.0, 1.0], "max_acc": [2.0, 2.0, 2.0], "radius": 0.15,
                                         "nominal_velocity": 1.0, "downwash": 2.0}},
            "world": [{"dimension": [-5.0, -5.0, 0.0, 5.0, 5.0, 2.5]}],
            "agents": [{"type": "crazyflie", "cid": i + 1, "start": s, "goal": [-s[0], -s[1], s[2]]} for i, s in enumerate(
                [[4.0, 0.0, 1], [3.0, 2.5, 1], [1.0, 4.0, 1], [-1.0, 4.0, 1], [-3.0, 2.5, 1], [-4.0, 0.0, 1], [-3.0, -2.5, 1],
                 [-1.0, -4.0, 1], [1.0, -4.0, 1], [3.0, -2.5, 1]])], "obstacles": []}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mission", nargs="?", default=None, help="mission JSON in the reference's format (default: forest10)")
    ap.add_argument("--out", default="gpurun_out/mission")
    ap.add_argument("--max-steps", type=int, default=600)            # multisim/max_planner_iteration
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--goal-mode", default="righthand", choices=["static", "righthand"])
    args = ap.parse_args()
    import torch
    from lsc_dr_planner_b200 import missions as MS, results as R, capi
    from lsc_dr_planner_b200.closed_loop import ClosedLoopSim
    mission = MS.load_mission(args.mission, args.dim, 1.0) if args.mission else MS.parse_mission(FOREST10, args.dim, 1.0)
    cfg = MS.launch_config(mission, M=5, dim=args.dim, comm_range=0.0)
    batch = MS.first_replan_batch(mission, cfg)
    batch.goal = mission.goal.copy()                                 # GoalMode static: fly to the desired goal
    n = mission.n_agents
    sim = ClosedLoopSim(batch, device=0, K=min(40, n - 1), generator=capi.GEN_LSC, goal_mode=args.goal_mode)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    writer = R.SimulationCsvWriter(args.out + "_simulation.csv", n, time_step=cfg.dt, record_time_step=0.1, dt=cfg.dt)
    positions = [batch.state[:, :3].copy()]
    step_times, t = [], 0.0
    for s in range(args.max_steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); sim.step(); ev1.record(); torch.cuda.synchronize()
        step_times.append(ev0.elapsed_time(ev1) * 1e-3 / n)          # the reference reports per-agent planning time
        writer.record(t, sim.traj_out.cpu().numpy(), step_times[-1])
        t += cfg.dt
        done = sim.max_goal_distance() < 0.1                            # (synchronises sim.state with the step just made)
        positions.append(sim.state[:, :3].cpu().numpy())
        if done:                            # plan/goal_threshold
            break
    dist, ratio = R.flight_metrics(np.stack(positions), mission.radius, mission.downwash)
    st = np.array(step_times)
    summary = R.MissionSummary(start_time="%.6f" % time.time(), total_flight_time=t, total_flight_distance=dist, safety_ratio_agent=ratio,
                               planning_time=(float(st.mean()), float(st.min()), float(st.max())),
                               stage_times=dict(traj_optimization=float(st.mean())),
                               mission_file_name=args.mission or "forest10 (built in)", world_file_name="(none)", planner_mode="LSC",
                               goal_mode=args.goal_mode, mapf_mode="none", communication_range=0.0, world_dimension=args.dim, M=cfg.M, dt=cfg.dt)
    R.append_summary_csv(args.out + "_summary.csv", summary)
    print(json.dumps({"agents": n, "replans": s + 1, "flight_time_s": t, "flight_distance_m": dist, "safety_ratio_agent": ratio,
                      "goal_mode": args.goal_mode, "qp_failures": sim.failed_total, "goal_distance_end": sim.max_goal_distance(),
                      "per_agent_planning_time_ms": float(st.mean() * 1e3), "files": [args.out + "_simulation.csv", args.out + "_summary.csv"]}))


if __name__ == "__main__":
    main()
