"""Synthetic N-agent replan batches.

Pure numpy: these build *inputs* only (states, goals, trajectories, neighbour lists,
synthetic half-spaces).  The half-spaces that follow the reference's real rule are
produced by the device assembly kernel (or, in tests, by the oracle) from the
trajectories generated here.

Layouts follow include/lscqp.h.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class PlannerConfig:
    """The subset of the reference's Param/Mission the QP reads:
    param.cpp:11,70-78,117 defaults, launch/simulation.launch:44-45,56,78,84-85 overrides."""
    M: int = 5
    n: int = 5
    phi: int = 3
    dim: int = 3
    dt: float = 0.2
    w_control: float = 0.01
    w_terminal: float = 1.0
    planner_mode: int = 1            # PlannerMode::LSC (sp_const.hpp:19-26)
    use_sfc: bool = False
    comm_range: float = 0.0
    world_min: tuple = (-10.0, -10.0, 0.0)
    world_max: tuple = (10.0, 10.0, 2.5)
    z_2d: float = 1.0
    max_obs: int = 40
    max_iter: int = 0            # 0 = library default (60)
    tol: float = 0.0             # 0 = library default (1e-11 on the mean complementarity gap)
    presolve: bool = True        # exact row pruning by bound propagation (lscqp_config.presolve)


def bernstein_from_poly(coef: np.ndarray, t0: float, t1: float) -> np.ndarray:
    """Quintic Bezier control points on [t0,t1] of the polynomial sum coef[j] t^j (coef [..., 6])."""
    # sample at 6 Chebyshev-free equispaced nodes and invert the Bernstein collocation matrix
    n = 5
    u = np.linspace(0.0, 1.0, n + 1)
    from math import comb
    Bm = np.array([[comb(n, i) * uu ** i * (1 - uu) ** (n - i) for i in range(n + 1)] for uu in u])
    t = t0 + u * (t1 - t0)
    V = np.stack([t ** j for j in range(6)], axis=0)             # [6 powers][6 nodes]
    vals = coef @ V                                               # [..., 6 nodes]
    return np.einsum("ij,...j->...i", np.linalg.inv(Bm), vals)


def braking_traj(p0, v0, a0, M: int, dt: float) -> np.ndarray:
    """A dynamically consistent 'previous solution': the single quintic with the given initial
    position/velocity/acceleration that comes to rest (zero velocity, acceleration and jerk) at
    the end of the horizon, cut into M Bezier segments.  It satisfies every equality of the
    reference QP (initial state, C2 continuity, terminal stop: last three control points equal).
    p0, v0, a0: [..., 3].  Returns float32 [..., M, 6, 3]."""
    p0, v0, a0 = (np.asarray(x, np.float64) for x in (p0, v0, a0))
    T = M * dt
    # p(t) = pT + al (t-T)^4 + be (t-T)^5 ;  p'(0) = -4 al T^3 + 5 be T^4 ; p''(0) = 12 al T^2 - 20 be T^3
    A = np.array([[-4 * T ** 3, 5 * T ** 4], [12 * T ** 2, -20 * T ** 3]])
    Ai = np.linalg.inv(A)
    al = Ai[0, 0] * v0 + Ai[0, 1] * a0
    be = Ai[1, 0] * v0 + Ai[1, 1] * a0
    pT = p0 - al * T ** 4 + be * T ** 5
    # expand in powers of t
    from math import comb
    coef = np.zeros(p0.shape + (6,))
    coef[..., 0] = pT
    for j in range(5):
        coef[..., j] += al * comb(4, j) * (-T) ** (4 - j)
    for j in range(6):
        coef[..., j] += be * comb(5, j) * (-T) ** (5 - j)
    segs = [bernstein_from_poly(coef, m * dt, (m + 1) * dt) for m in range(M)]   # each [..., 3, 6]
    out = np.stack(segs, axis=-3)                                 # [..., M, 3, 6]
    out = np.swapaxes(out, -1, -2)                                # [..., M, 6, 3]
    # exact equality structure in float: junction points shared, terminal stop
    out = out.astype(np.float32)
    out[..., M - 1, 3, :] = out[..., M - 1, 5, :]
    out[..., M - 1, 4, :] = out[..., M - 1, 5, :]
    return out


def traj_limits_ok(traj: np.ndarray, dt: float, vmax: float, amax: float) -> np.ndarray:
    """Velocity / acceleration control-point bounds of the reference QP (traj_optimizer.cpp:440-474)."""
    t = traj.astype(np.float64)
    v = 5.0 / dt * (t[..., 1:, :] - t[..., :-1, :])
    a = 20.0 / dt ** 2 * (t[..., 2:, :] - 2 * t[..., 1:-1, :] + t[..., :-2, :])
    return (np.abs(v).max(axis=(-1, -2, -3)) <= vmax) & (np.abs(a).max(axis=(-1, -2, -3)) <= amax)


@dataclass
class Batch:
    cfg: PlannerConfig
    state: np.ndarray          # [N,9]  f32  pos, vel, acc
    goal: np.ndarray           # [N,3]  f32
    limits: np.ndarray         # [N,8]  f64  vmax[3], amax[3], radius, nominal_velocity
    next_waypoint: np.ndarray  # [N,3]  f32
    agent_meta: np.ndarray     # [N,2]  f64  radius, downwash
    own_traj: np.ndarray       # [N,M,6,3] f32 initial_traj
    obs_offsets: np.ndarray    # [N+1] i32
    obs_index: np.ndarray      # [sumK] i32 neighbour agent ids
    sfc: np.ndarray | None = None   # [N,M,6] f32

    @property
    def n_agents(self) -> int:
        return self.state.shape[0]

    def obs_traj(self) -> np.ndarray:
        return np.ascontiguousarray(self.own_traj[self.obs_index])

    def obs_meta(self) -> np.ndarray:
        m = np.zeros((self.obs_index.size, 4), np.float32)
        m[:, 0] = self.agent_meta[self.obs_index, 0]
        m[:, 1] = self.agent_meta[self.obs_index, 1]
        m[:, 2] = 1.0          # ObstacleType::AGENT
        return m

    def obs_goal(self) -> np.ndarray:
        return np.ascontiguousarray(self.goal[self.obs_index])

    def obs_position(self) -> np.ndarray:
        return np.ascontiguousarray(self.state[self.obs_index, :3])


def _sample_positions(rng, n, cell, jitter, zlo, zhi, dim):
    """jittered square grid (cell size `cell`, xy jitter +-`jitter`): minimum horizontal separation
    cell - 2*jitter by construction, O(n)."""
    side = int(np.ceil(np.sqrt(n)))
    ij = np.stack(np.meshgrid(np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 2)
    ij = ij[rng.permutation(side * side)[:n]]
    xy = (ij - (side - 1) / 2.0) * cell + rng.uniform(-jitter, jitter, (n, 2))
    z = rng.uniform(zlo, zhi, n) if dim == 3 else np.full(n, 1.0)
    return np.column_stack([xy, z]), side * cell / 2.0


def make_forest_batch(n_agents: int, K: int = 40, seed: int = 20260001,
                      cfg: PlannerConfig | None = None, moving: bool = True) -> Batch:
    """Config 2 style: agents in a random forest, K nearest neighbours each, goals antipodal,
    every agent carrying a braking 'previous solution' as initial_traj."""
    cfg = cfg or PlannerConfig()
    rng = np.random.default_rng(seed)
    M, dt = cfg.M, cfg.dt
    radius, downwash, vmax, amax = 0.15, 2.0, 1.0, 2.0
    # horizontal separation >= 1.3 m: the braking trajectories' hulls stay >= 2r apart
    pos, half = _sample_positions(rng, n_agents, 2.0, 0.35, 0.5, 2.0, cfg.dim)
    cfg.world_min = (-half - 2.0, -half - 2.0, 0.0)
    cfg.world_max = (half + 2.0, half + 2.0, 2.5)
    vel = np.zeros((n_agents, 3)); acc = np.zeros((n_agents, 3))
    if moving:
        vel = rng.uniform(-0.45, 0.45, (n_agents, 3)); acc = rng.uniform(-0.5, 0.5, (n_agents, 3))
        if cfg.dim == 2:
            vel[:, 2] = 0; acc[:, 2] = 0
    traj = braking_traj(pos, vel, acc, M, dt)
    ok = traj_limits_ok(traj, dt, vmax, amax)
    vel[~ok] = 0; acc[~ok] = 0
    traj = braking_traj(pos, vel, acc, M, dt)
    # keep z inside the world
    goal = -pos.copy(); goal[:, 2] = pos[:, 2]
    if cfg.dim == 2:
        goal[:, 2] = cfg.z_2d
    from scipy.spatial import cKDTree
    k = min(K, n_agents - 1)
    tree = cKDTree(pos)
    _, idx = tree.query(pos, k=k + 1)
    idx = np.asarray(idx).reshape(n_agents, k + 1)[:, 1:].astype(np.int32)
    state = np.concatenate([traj[:, 0, 0, :],
                            (5.0 / dt) * (traj[:, 0, 1, :] - traj[:, 0, 0, :]),
                            (20.0 / dt ** 2) * (traj[:, 0, 2, :] - 2 * traj[:, 0, 1, :] + traj[:, 0, 0, :])],
                           axis=1).astype(np.float32)
    limits = np.tile(np.array([vmax] * 3 + [amax] * 3 + [radius, 1.0]), (n_agents, 1))
    meta = np.zeros((n_agents, 2), np.float64); meta[:, 0] = radius; meta[:, 1] = downwash
    return Batch(cfg, state, goal.astype(np.float32), limits, goal.astype(np.float32), meta,
                 traj, (np.arange(n_agents + 1) * k).astype(np.int32), idx.reshape(-1))


def make_synthetic_planes(batch: Batch, K: int = 40, seed: int = 20260004):
    """Config 4: random half-spaces with a strictly feasible point (the initial trajectory).
    For each (agent, oi, m): unit normal n, rhs_i = n . c_{m,i} - margin - jitter_i with
    margin ~ U(0.02, 0.5), jitter ~ U(-0.01, 0.01).  Returns (offsets[N+1] i32,
    normals [N*K, M, 3] f64, rhs [N*K, M, 6] f64): rows read n . c >= rhs."""
    rng = np.random.default_rng(seed)
    N, M = batch.n_agents, batch.cfg.M
    nrm = rng.normal(size=(N, K, M, 3))
    if batch.cfg.dim == 2:
        nrm[..., 2] = 0
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    nrm = nrm.astype(np.float32).astype(np.float64)       # the reference stores normals as float
    c = batch.own_traj.astype(np.float64)                 # [N,M,6,3]
    dots = np.einsum("nkmd,nmid->nkmi", nrm, c)
    margin = rng.uniform(0.02, 0.5, (N, K, M, 1))
    jitter = rng.uniform(-0.01, 0.01, (N, K, M, 6))
    rhs = dots - margin - jitter
    offsets = (np.arange(N + 1) * K).astype(np.int32)
    return offsets, nrm.reshape(N * K, M, 3), rhs.reshape(N * K, M, 6)
