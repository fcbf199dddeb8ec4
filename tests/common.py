"""Shared helpers for the parity tests: oracle-side models of a workloads.Batch."""
from __future__ import annotations

import numpy as np

from lsc_dr_planner_b200 import workloads as W
from oracle import oracle as orc


def oracle_config(cfg: W.PlannerConfig) -> orc.Config:
    return orc.Config(M=cfg.M, n=cfg.n, phi=cfg.phi, dim=cfg.dim, dt=cfg.dt, w_control=cfg.w_control,
                      w_terminal=cfg.w_terminal, planner_mode=cfg.planner_mode, use_sfc=cfg.use_sfc,
                      comm_range=cfg.comm_range, world_min=cfg.world_min, world_max=cfg.world_max, z_2d=cfg.z_2d)


def oracle_agent(batch: W.Batch, a: int) -> orc.Agent:
    lim = batch.limits[a]
    return orc.Agent(batch.state[a, :3], batch.state[a, 3:6], batch.state[a, 6:9], batch.goal[a],
                     next_waypoint=batch.next_waypoint[a], max_vel=tuple(lim[:3]), max_acc=tuple(lim[3:6]),
                     radius=float(batch.agent_meta[a, 0]), nominal_velocity=float(lim[7]),
                     downwash=float(batch.agent_meta[a, 1]))


def oracle_lsc(batch: W.Batch, a: int, generator: int):
    """reference-rule LSC records of agent a: (point, normal, d)"""
    cfgo = oracle_config(batch.cfg)
    sl = slice(batch.obs_offsets[a], batch.obs_offsets[a + 1])
    return orc.generate_lsc(cfgo, generator, oracle_agent(batch, a), batch.own_traj[a], batch.obs_traj()[sl],
                            batch.obs_meta()[sl, 0], batch.obs_meta()[sl, 1], batch.obs_goal()[sl],
                            batch.obs_position()[sl])


def oracle_planes(batch: W.Batch, agents, generator: int):
    """packed planes (offsets, normals, rhs) for a list of agents, by the oracle"""
    cfgo = oracle_config(batch.cfg)
    normals, rhs, off = [], [], [0]
    for a in agents:
        pt, nr, d = oracle_lsc(batch, a, generator)
        n_, r_ = orc.pack_planes(cfgo, pt, nr, d)
        normals.append(n_); rhs.append(r_); off.append(off[-1] + n_.shape[0])
    M = batch.cfg.M
    normals = np.concatenate(normals) if normals else np.zeros((0, M, 3))
    rhs = np.concatenate(rhs) if rhs else np.zeros((0, M, 6))
    return np.array(off, np.int32), np.ascontiguousarray(normals), np.ascontiguousarray(rhs)


def oracle_qp_from_planes(batch: W.Batch, a: int, normals, rhs, sfc=None) -> orc.QP:
    """restated populatebyrow model for agent a whose LSCs are the packed planes (point = 0, d = rhs)"""
    cfgo = oracle_config(batch.cfg)
    nr = np.repeat(np.asarray(normals, np.float32)[:, :, None, :], 6, axis=2)
    pt = np.zeros_like(nr)
    return orc.qp_build(cfgo, oracle_agent(batch, a), pt, nr, np.asarray(rhs, np.float64), sfc)


def oracle_solution(qp: orc.QP):
    """(x, ok): HiGHS solution polished to ~1e-12; ok False when HiGHS failed or polish was refused"""
    sol = orc.solve_highs(qp)
    if sol.status != "Optimal":
        sol = orc.solve_dense_ipm(qp)       # HiGHS' QP solver errors out on some dense comm-range models
        if sol.status != "Optimal":
            return sol.x, False
        return orc.polish(qp, sol, dual_tol=1e-12)
    return orc.polish(qp, sol)


def near_goals(batch: W.Batch, seed: int = 5, spread: float = 0.5) -> None:
    """current_goal_point close to the end of the previous solution (what goalPlanning produces in closed loop)"""
    rng = np.random.default_rng(seed)
    g = batch.own_traj[:, -1, -1, :] + rng.uniform(-spread, spread, (batch.n_agents, 3)).astype(np.float32)
    if batch.cfg.dim == 2:
        g[:, 2] = batch.cfg.z_2d
    else:
        g[:, 2] = np.clip(g[:, 2], 0.3, 2.2)
    batch.goal = g.astype(np.float32)
