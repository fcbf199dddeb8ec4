"""Result files in the reference's formats -- harness-level plumbing, no hot-path arithmetic.

  * per-step CSV   MultiSyncSimulator::saveSimulationResultAsCSV (src/multi_sync_simulator.cpp:586-651): one block of
                   "id,t,px,py,pz,vx,vy,vz,ax,ay,az,planning_time" per agent on every line, sampled every
                   multisim/record_time_step (0.1 s) inside each replanning period from the agent's desired trajectory.
  * summary CSV    saveSummarizedResultAsCSV (:653-709): one row per mission with the 28 columns of log/summary_*.csv.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from math import comb

import numpy as np

SUMMARY_COLUMNS = ("start_time,total_flight_time,total_flight_distance,safety_ratio_agent,safety_ratio_obs,"
                   "vel_excess_ratio,acc_excess_ratio,mapf_time_average,mapf_time_min,mapf_time_max,"
                   "planning_time_average,planning_time_min,planning_time_max,initial_traj_planning_time,"
                   "obstacle_prediction_time,goal_planning_time,lsc_generation_time,sfc_generation_time,"
                   "traj_optimization_time,mission_file_name,world_file_name,planner_mode,goal_mode,mapf_mode,"
                   "communication_range,world_dimension,M,dt").split(",")
AGENT_COLUMNS = "id,t,px,py,pz,vx,vy,vz,ax,ay,az,planning_time"
SP_INFINITY = 1e9                                      # include/sp_const.hpp:5


def fmt(x) -> str:
    """std::ostream's default formatting of a double / int (6 significant digits, %g)"""
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    if isinstance(x, str):
        return x
    return "%g" % float(x)


def states_at(traj: np.ndarray, dt: float, time: float) -> np.ndarray:
    """Trajectory::getStateAt (src/trajectory.cpp:111-199) for a batch of float trajectories [N, M, 6, 3]:
    position, velocity, acceleration at `time` -> [N, 9] float32 (the derivative trajectories are formed in float)."""
    traj = np.asarray(traj, np.float32)
    N, M = traj.shape[0], traj.shape[1]
    out = np.zeros((N, 9), np.float32)
    d1 = ((traj[:, :, 1:, :] - traj[:, :, :-1, :]) * np.float32(5.0 / dt)).astype(np.float32)
    d2 = ((d1[:, :, 1:, :] - d1[:, :, :-1, :]) * np.float32(4.0 / dt)).astype(np.float32)
    for k, (cps, deg) in enumerate(((traj, 5), (d1, 4), (d2, 3))):
        if time < 0:
            continue
        m = min(int(np.floor(time / dt + 1e-12)), M)
        seg_end = (m + 1) * dt
        if m >= M:
            if time < M * dt + 1e-5:
                m, t_norm = M - 1, 1.0
            else:
                continue
        else:
            t_norm = 1.0 - (seg_end - time) / dt
        acc = np.zeros((N, 3), np.float32)
        for i in range(deg + 1):
            b = comb(deg, i) * t_norm ** i * (1.0 - t_norm) ** (deg - i)
            acc = (acc + cps[:, m, i, :] * np.float32(b)).astype(np.float32)
        out[:, 3 * k:3 * k + 3] = acc
    return out


class SimulationCsvWriter:
    """saveSimulationResultAsCSV: call `record(t, desired_traj, planning_time)` once per replanning period"""

    def __init__(self, path: str, n_agents: int, time_step: float = 0.2, record_time_step: float = 0.1, dt: float = 0.2):
        self.path, self.n, self.time_step, self.record_time_step, self.dt = path, n_agents, time_step, record_time_step, dt
        with open(path, "w") as f:
            f.write(",".join([AGENT_COLUMNS] * n_agents) + "\n")

    def record(self, t: float, desired_traj: np.ndarray, planning_time: np.ndarray | float = 0.0) -> None:
        pt = np.broadcast_to(np.asarray(planning_time, np.float64), (self.n,))
        future, lines = 0.0, []
        while future < self.time_step:
            st = states_at(desired_traj, self.dt, future)
            lines.append(",".join(",".join([str(qi), fmt(t)] + [fmt(v) for v in st[qi]] + [fmt(pt[qi])]) for qi in range(self.n)))
            future += self.record_time_step
            t += self.record_time_step
        with open(self.path, "a") as f:
            f.write("\n".join(lines) + "\n")


@dataclass
class MissionSummary:
    """the quantities of one row of log/summary_*.csv; times in seconds (Timer averages over agents and replans)"""
    start_time: str = "0"
    total_flight_time: float = 0.0
    total_flight_distance: float = 0.0
    safety_ratio_agent: float = SP_INFINITY
    safety_ratio_obs: float = SP_INFINITY
    vel_excess_ratio: float = 0.0
    acc_excess_ratio: float = 0.0
    mapf_time: tuple = (0.0, 0.0, 0.0)                  # average, min, max
    planning_time: tuple = (0.0, 0.0, 0.0)
    stage_times: dict = field(default_factory=dict)    # initial_traj_planning, obstacle_prediction, goal_planning, lsc_generation, sfc_generation, traj_optimization
    mission_file_name: str = ""
    world_file_name: str = ""
    planner_mode: str = "LSC"
    goal_mode: str = "grid_based_planner"
    mapf_mode: str = "pibt"
    communication_range: float = 3.0
    world_dimension: int = 2
    M: int = 10
    dt: float = 0.2

    def row(self) -> list:
        st = self.stage_times
        return [self.start_time, self.total_flight_time, self.total_flight_distance, self.safety_ratio_agent, self.safety_ratio_obs,
                self.vel_excess_ratio, self.acc_excess_ratio, *self.mapf_time, *self.planning_time,
                st.get("initial_traj_planning", 0.0), st.get("obstacle_prediction", 0.0), st.get("goal_planning", 0.0),
                st.get("lsc_generation", 0.0), st.get("sfc_generation", 0.0), st.get("traj_optimization", 0.0),
                self.mission_file_name, self.world_file_name, self.planner_mode, self.goal_mode, self.mapf_mode,
                self.communication_range, self.world_dimension, self.M, self.dt]


def append_summary_csv(path: str, summary: MissionSummary) -> None:
    """saveSummarizedResultAsCSV: the header is written only into a new / empty file, rows are appended"""
    new = not os.path.exists(path) or os.path.getsize(path) == 0
    with open(path, "a") as f:
        if new:
            f.write(",".join(SUMMARY_COLUMNS) + "\n")
        f.write(",".join(fmt(v) for v in summary.row()) + "\n")


def flight_metrics(positions: np.ndarray, radius: np.ndarray, downwash: np.ndarray) -> tuple[float, float]:
    """total flight distance and minimum agent safety ratio of a recorded run: positions [T, N, 3].
    Safety ratio = downwash-scaled centre distance / (r_i + r_j) minimised over pairs and samples
    (src/multi_sync_simulator.cpp:541-583, the quantity summary_*.csv reports as safety_ratio_agent)."""
    p = np.asarray(positions, np.float64)
    dist = float(np.linalg.norm(np.diff(p, axis=0), axis=2).sum())
    worst = SP_INFINITY
    r = np.asarray(radius, np.float64); dw = np.asarray(downwash, np.float64)
    for t in range(p.shape[0]):
        q = p[t].copy()
        d = q[:, None, :] - q[None, :, :]
        dwp = (dw[:, None] * r[:, None] + dw[None, :] * r[None, :]) / (r[:, None] + r[None, :])
        d[..., 2] /= dwp
        ratio = np.linalg.norm(d, axis=2) / (r[:, None] + r[None, :])
        np.fill_diagonal(ratio, np.inf)
        worst = min(worst, float(ratio.min()))
    return dist, worst
