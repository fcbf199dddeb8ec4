"""Mission / world file readers and mission-driven replan batches.

Reads the reference's own input formats -- harness-level plumbing, no arithmetic of the hot path:
  * mission JSON, Mission::readMissionFile (src/mission.cpp:94-260): "quadrotors" (per type: max_vel[3], max_acc[3],
    radius, downwash, nominal_velocity), "world" (one element, "dimension": min xyz, max xyz), "agents"
    (type, optional cid, start, goal, optional downwash / nominal_velocity overrides); in 2-D the z of start and
    goal is replaced by world_z_2d (:166-169, :184-187).  Scripted "obstacles" are not supported (none of the shipped
    missions has any) and are rejected.
  * world CSV (src/map_manager.cpp:262-305): rows "cx,cy,cz,sx,sy,sz" = axis-aligned boxes by centre and size.
"""
from __future__ import annotations

import json
from dataclasses import dataclass

import numpy as np

from .workloads import Batch, PlannerConfig


@dataclass
class Mission:
    world_min: tuple
    world_max: tuple
    start: np.ndarray            # [N,3] f32
    goal: np.ndarray             # [N,3] f32  desired_goal_point
    max_vel: np.ndarray          # [N,3] f64
    max_acc: np.ndarray          # [N,3] f64
    radius: np.ndarray           # [N]
    downwash: np.ndarray         # [N]
    nominal_velocity: np.ndarray  # [N]
    cid: np.ndarray              # [N] i32

    @property
    def n_agents(self) -> int:
        return self.start.shape[0]


def parse_mission(doc: dict, world_dimension: int = 3, world_z_2d: float = 1.0) -> Mission:
    """Mission::readMissionFile on an already parsed JSON document"""
    world = doc["world"]
    if len(world) != 1:
        raise ValueError("[Mission] World must have one element")                 # mission.cpp:106-109
    dim = [np.float32(v) for v in world[0]["dimension"]]
    quads = doc["quadrotors"]
    if doc.get("obstacles"):
        raise ValueError("scripted obstacles are outside the batched agent-QP path")
    agents = doc["agents"]
    n = len(agents)
    out = Mission(tuple(float(v) for v in dim[:3]), tuple(float(v) for v in dim[3:]), np.zeros((n, 3), np.float32),
                  np.zeros((n, 3), np.float32), np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n), np.zeros(n), np.zeros(n),
                  np.zeros(n, np.int32))
    for qi, a in enumerate(agents):
        if "type" not in a:
            raise ValueError("[Mission] Agent must have type element")
        q = quads[a["type"]]
        out.max_vel[qi] = q["max_vel"]; out.max_acc[qi] = q["max_acc"]
        out.radius[qi] = q["radius"]; out.downwash[qi] = q["downwash"]; out.nominal_velocity[qi] = q["nominal_velocity"]
        out.cid[qi] = a.get("cid", qi)
        for key, dst in (("start", out.start), ("goal", out.goal)):
            if key not in a:
                raise ValueError(f"[Mission] Agent must have {key} element")
            p = a[key]
            dst[qi] = (p[0], p[1], world_z_2d if world_dimension == 2 else p[2])
        if "size" in a:                                                          # (sic) mission.cpp:191-193
            out.radius[qi] = a["radius"]
        if "downwash" in a:
            out.downwash[qi] = a["downwash"]
        if "nominal_velocity" in a:
            out.nominal_velocity[qi] = a["nominal_velocity"]
    return out


def load_mission(path: str, world_dimension: int = 3, world_z_2d: float = 1.0) -> Mission:
    with open(path) as f:
        return parse_mission(json.load(f), world_dimension, world_z_2d)


def load_world_csv(path: str) -> np.ndarray:
    """boxes [n,6] = centre xyz, size xyz (map_manager.cpp:262-305)"""
    rows = [[float(v) for v in line.split(",")[:6]] for line in open(path) if line.strip()]
    return np.asarray(rows, np.float64).reshape(-1, 6)


def neighbours_linf(position: np.ndarray, comm_range: float) -> tuple[np.ndarray, np.ndarray]:
    """MultiSyncSimulator::broadcastMsgs (src/multi_sync_simulator.cpp:305-352): agent qj is an obstacle of qi
    unless qi == qj or (communication_range > 0 and the Chebyshev distance of the current positions exceeds it).
    Returns CSR (offsets[N+1], index[sum K]) in agent order."""
    n = position.shape[0]
    p = position.astype(np.float32).astype(np.float64)
    off, idx = [0], []
    for qi in range(n):
        d = np.abs(p - p[qi]).max(axis=1)
        for qj in range(n):
            if qj == qi or (comm_range > 0 and d[qj] > comm_range):
                continue
            idx.append(qj)
        off.append(len(idx))
    return np.asarray(off, np.int32), np.asarray(idx, np.int32)


def launch_config(mission: Mission, M: int = 10, dim: int = 2, comm_range: float = 3.0, z_2d: float = 1.0) -> PlannerConfig:
    """the shipped launch settings (launch/simulation.launch:44-85): LSC mode, M=10, n=5, dt=0.2, 2-D,
    control_input_weight 0.01, terminal_weight 1, communication range 3.  SFC boxes are an input of this path
    (octomap / dynamicEDT3D are out of scope), so use_sfc is off unless the caller supplies boxes."""
    return PlannerConfig(M=M, dim=dim, dt=0.2, w_control=0.01, w_terminal=1.0, planner_mode=1, use_sfc=False,
                         comm_range=comm_range, world_min=mission.world_min, world_max=mission.world_max, z_2d=z_2d,
                         max_obs=min(40, max(1, mission.n_agents - 1)))


def first_replan_batch(mission: Mission, cfg: PlannerConfig, waypoint_step: float = 0.5) -> Batch:
    """Inputs of the first replan (planner_seq == 1): every agent at rest at its start point, own and neighbours'
    trajectories constant at the current position (traj_planner.cpp:276-279, 400-401 with zero velocity),
    current_goal_point = next_waypoint = start (agent_manager.cpp:9-10) -- then the waypoint layer moves
    next_waypoint one grid cell along the path (multi_sync_simulator.cpp:219-220).  The grid planner is outside this
    path; its first waypoint is stood in for by the point `waypoint_step` (grid/resolution 0.5) from the start
    towards the desired goal."""
    n, M = mission.n_agents, cfg.M
    state = np.zeros((n, 9), np.float32); state[:, :3] = mission.start
    own = np.repeat(np.repeat(mission.start[:, None, None, :], M, 1), 6, 2).astype(np.float32)
    to_goal = mission.goal.astype(np.float64) - mission.start.astype(np.float64)
    dist = np.linalg.norm(to_goal, axis=1, keepdims=True)
    wp = (mission.start + to_goal / np.maximum(dist, 1e-9) * np.minimum(dist, waypoint_step)).astype(np.float32)
    limits = np.concatenate([mission.max_vel, mission.max_acc, mission.radius[:, None], mission.nominal_velocity[:, None]], 1)
    meta = np.stack([mission.radius, mission.downwash], 1)
    off, idx = neighbours_linf(mission.start, cfg.comm_range)
    return Batch(cfg, state, mission.start.copy(), np.ascontiguousarray(limits), wp, np.ascontiguousarray(meta), own, off, idx)
