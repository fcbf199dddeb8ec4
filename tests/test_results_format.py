"""Result-file formats of the reference (saveSimulationResultAsCSV / saveSummarizedResultAsCSV,
src/multi_sync_simulator.cpp:586-709) and the host-side state sampling they need."""
import numpy as np

from common import oracle_config
from lsc_dr_planner_b200 import results as R
from lsc_dr_planner_b200 import workloads as W
from oracle import oracle as orc

# header and first data row of the reference's log/summary_LSC_10agents.csv
REF_HEADER = ("start_time,total_flight_time,total_flight_distance,safety_ratio_agent,safety_ratio_obs,vel_excess_ratio,"
              "acc_excess_ratio,mapf_time_average,mapf_time_min,mapf_time_max,planning_time_average,planning_time_min,"
              "planning_time_max,initial_traj_planning_time,obstacle_prediction_time,goal_planning_time,lsc_generation_time,"
              "sfc_generation_time,traj_optimization_time,mission_file_name,world_file_name,planner_mode,goal_mode,mapf_mode,"
              "communication_range,world_dimension,M,dt")
REF_ROW = ("1663743693.650981,15.8,103.163,1.02089,1e+09,0,0,3.38685e-05,0,0.000586999,0.00825623,0.00462922,0.0271518,"
           "1.77482e-06,1.4475e-05,0.000181179,9.43503e-05,0.00132031,0.00663744,missions/forest10/forest10_10.json,"
           "world/forest/forest10.csv,LSC,grid_based_planner,pibt,3,2,10,0.2")


def test_summary_csv_reproduces_the_reference_row(tmp_path):
    s = R.MissionSummary(start_time="1663743693.650981", total_flight_time=15.8, total_flight_distance=103.163,
                         safety_ratio_agent=1.02089, mapf_time=(3.38685e-05, 0, 0.000586999),
                         planning_time=(0.00825623, 0.00462922, 0.0271518),
                         stage_times=dict(initial_traj_planning=1.77482e-06, obstacle_prediction=1.4475e-05, goal_planning=0.000181179,
                                          lsc_generation=9.43503e-05, sfc_generation=0.00132031, traj_optimization=0.00663744),
                         mission_file_name="missions/forest10/forest10_10.json", world_file_name="world/forest/forest10.csv")
    p = tmp_path / "summary.csv"
    R.append_summary_csv(str(p), s); R.append_summary_csv(str(p), s)
    lines = p.read_text().splitlines()
    assert lines[0] == REF_HEADER and lines[1] == REF_ROW and lines[2] == REF_ROW and len(lines) == 3


def test_simulation_csv_layout_and_state_sampling(tmp_path):
    cfg = W.PlannerConfig(M=5, dim=3)
    batch = W.make_forest_batch(6, K=2, cfg=cfg)
    cfgo = oracle_config(cfg)
    for t in (0.0, 0.1, 0.2, 0.37, 1.0):
        want = np.stack([orc.get_state_at(cfgo, batch.own_traj[a], t) for a in range(6)])
        got = R.states_at(batch.own_traj, cfg.dt, t)
        assert np.abs(got - want).max() <= 2e-6 * max(1.0, np.abs(want).max()), (t, np.abs(got - want).max())
    p = tmp_path / "sim.csv"
    w = R.SimulationCsvWriter(str(p), 6)
    w.record(0.0, batch.own_traj, planning_time=0.001)
    w.record(0.2, batch.own_traj, planning_time=np.full(6, 0.002))
    lines = p.read_text().splitlines()
    assert lines[0] == ",".join([R.AGENT_COLUMNS] * 6) and len(lines) == 1 + 2 * 2          # two samples per 0.2 s period
    first = lines[1].split(",")
    assert len(first) == 12 * 6 and first[0] == "0" and first[1] == "0" and first[12] == "1"
    assert abs(float(first[2]) - float(batch.own_traj[0, 0, 0, 0])) < 1e-5 and first[11] == "0.001"
    assert lines[3].split(",")[1] == "0.2" and lines[4].split(",")[1].startswith("0.3")


def test_flight_metrics():
    pos = np.zeros((3, 2, 3)); pos[:, 1, 0] = [1.0, 0.8, 0.6]; pos[:, 1, 2] = [0.0, 0.0, 0.6]
    dist, ratio = R.flight_metrics(pos, np.array([0.15, 0.15]), np.array([2.0, 2.0]))
    assert abs(dist - (0.2 + np.hypot(0.2, 0.6))) < 1e-12
    assert abs(ratio - np.hypot(0.6, 0.3) / 0.3) < 1e-12


def test_simulation_csv_first_row_matches_the_reference_log():
    """the reference's own golden log (log/simulation_1663743693.650981_LSC_10agents.csv: forest10, 2-D, world/z_2d = 0.6):
    header identical, and the first data row -- every agent at rest at its start point at t = 0 -- identical field by field
    except the planning-time column (a wall-clock measurement)"""
    import json
    from test_missions import FOREST10
    from lsc_dr_planner_b200 import missions as MS
    ref_header = ",".join(["id,t,px,py,pz,vx,vy,vz,ax,ay,az,planning_time"] * 10)
    ref_first = ("0,0,4,0,0.6,0,0,0,0,0,0,1,0,3,2.5,0.6,0,0,0,0,0,0,2,0,1,4,0.6,0,0,0,0,0,0,3,0,-1,4,0.6,0,0,0,0,0,0,"
                 "4,0,-3,2.5,0.6,0,0,0,0,0,0,5,0,-4,0,0.6,0,0,0,0,0,0,6,0,-3,-2.5,0.6,0,0,0,0,0,0,7,0,-1,-4,0.6,0,0,0,0,0,0,"
                 "8,0,1,-4,0.6,0,0,0,0,0,0,9,0,3,-2.5,0.6,0,0,0,0,0,0")
    mission = MS.parse_mission(json.loads(json.dumps(FOREST10)), world_dimension=2, world_z_2d=0.6)
    cfg = MS.launch_config(mission, z_2d=0.6)
    batch = MS.first_replan_batch(mission, cfg)
    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "sim.csv")
        w = R.SimulationCsvWriter(p, mission.n_agents, time_step=0.2, record_time_step=0.1, dt=cfg.dt)
        w.record(0.0, batch.own_traj, planning_time=0.0271518)
        lines = open(p).read().splitlines()
    assert lines[0] == ref_header
    got = lines[1].split(",")
    assert ",".join(",".join(got[i * 12:i * 12 + 11]) for i in range(10)) == ref_first
    assert got[11] == "0.0271518"                      # same %g formatting as the reference's stream output
