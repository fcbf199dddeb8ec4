"""BatchPlanner -- the batched dispatch that replaces the reference's serial per-agent loop
(MultiSyncSimulator::plan, src/multi_sync_simulator.cpp:354-374 -> AgentManager::plan ->
TrajPlanner::planImpl's constructLSC + trajOptimization, src/traj_planner.cpp:117-139).

Host-side orchestration only: all arithmetic happens in liblscqp.so's CUDA kernels through the
C ABI (capi.LscQp).  torch is used for device memory, streams and pinned buffers.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .workloads import Batch, PlannerConfig


class DeviceBatch:
    """A Batch resident in HBM plus the scratch the replan step needs."""

    def __init__(self, batch: Batch, device: int):
        import torch
        dev = torch.device("cuda", device)
        cfg = batch.cfg
        N, M = batch.n_agents, cfg.M
        self.n = N
        self.sum_k = int(batch.obs_offsets[-1])
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.state = t(batch.state)
        self.goal = t(batch.goal)
        self.limits = t(batch.limits)
        self.own_traj = t(batch.own_traj)
        self.agent_meta = t(batch.agent_meta)
        self.obs_offsets = t(batch.obs_offsets.astype(np.int32))
        self.obs_index = t(batch.obs_index.astype(np.int32)) if batch.obs_index.size else torch.zeros(1, dtype=torch.int32, device=dev)
        self.sfc = t(batch.sfc) if batch.sfc is not None else None
        self.next_waypoint = t(batch.next_waypoint) if cfg.comm_range > 0 else None
        self.waypoint = t(batch.next_waypoint)                      # GoalOptimizer input (any configuration)
        self.goal_new = torch.empty((N, 3), dtype=torch.float32, device=dev)
        self.goal_status = torch.empty((N,), dtype=torch.int32, device=dev)
        sk = max(self.sum_k, 1)
        self.obs_traj = torch.empty((sk, M, 6, 3), dtype=torch.float32, device=dev)
        self.obs_meta = torch.empty((sk, 4), dtype=torch.float32, device=dev)
        self.obs_goal = torch.empty((sk, 3), dtype=torch.float32, device=dev)
        self.obs_position = torch.empty((sk, 3), dtype=torch.float32, device=dev)
        self.normals = torch.empty((sk, M, 3), dtype=torch.float64, device=dev)
        self.rhs = torch.empty((sk, M, 6), dtype=torch.float64, device=dev)
        self.ctrl = torch.empty((N, cfg.dim * M * 6), dtype=torch.float64, device=dev)
        self.cost = torch.empty((N,), dtype=torch.float64, device=dev)
        self.status = torch.empty((N,), dtype=torch.int32, device=dev)
        self.iters = torch.empty((N,), dtype=torch.int32, device=dev)
        self.kkt = torch.empty((N, 4), dtype=torch.float64, device=dev)


class BatchPlanner:
    def __init__(self, cfg: PlannerConfig, device: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise capi.LscqpError("BatchPlanner needs a CUDA device; there is no CPU fallback")
        torch.cuda.set_device(device)
        self.cfg = cfg
        self.device = device
        self.qp = capi.LscQp(cfg, device)

    # ------------------------------------------------------------------ device-resident path
    def upload(self, batch: Batch) -> DeviceBatch:
        return DeviceBatch(batch, self.device)

    def assemble_device(self, d: DeviceBatch, generator: int = capi.GEN_LSC, stream: int = 0):
        """gather neighbours + LSC assembly (constructLSC for every agent): the reference's planes for every
        (obstacle, segment), materialised obstacle copies (the layout lscqp_assemble_lsc_batch documents)"""
        self.qp.gather_obstacles(d.sum_k, d.obs_index, d.own_traj, d.agent_meta, d.goal, d.state,
                                 d.obs_traj, d.obs_meta, d.obs_goal, d.obs_position, stream)
        self.qp.assemble_lsc_batch(generator, d.n, d.own_traj, d.agent_meta, d.goal, d.obs_offsets, d.obs_traj,
                                   d.obs_meta, d.obs_goal, d.obs_position, d.normals, d.rhs, stream)

    def assemble_fused_device(self, d: DeviceBatch, generator: int = capi.GEN_LSC, stream: int = 0, prune: bool | None = None):
        """the replan path's assembly: obstacles read in place through the neighbour ids (no gathered copies) and, with
        presolve on, (obstacle, segment) pairs that provably cannot bind dropped before the hull enumeration (exact)"""
        prune = bool(int(self.cfg.presolve) & 1) if prune is None else prune
        self.qp.assemble_lsc_fused(generator, prune, d.n, d.own_traj, d.agent_meta, d.goal, d.state, d.limits, d.obs_offsets,
                                   d.obs_index, d.own_traj, d.agent_meta, d.goal, d.state, d.normals, d.rhs, stream)

    def solve_device(self, d: DeviceBatch, want_kkt: bool = False, dual=None, stream: int = 0, warm: bool = True):
        """trajOptimization for every agent (initial_traj = the batch's own_traj as the solver's starting point)"""
        self.qp.solve_batch(d.n, d.state, d.goal, d.limits, d.sfc, d.obs_offsets, d.normals, d.rhs,
                            d.ctrl, d.cost, d.status, d.iters, d.kkt if want_kkt else None, dual, stream,
                            initial_traj=d.own_traj if warm else None, next_waypoint=d.next_waypoint)

    def replan_device(self, d: DeviceBatch, generator: int = capi.GEN_LSC, stream: int = 0):
        self.assemble_fused_device(d, generator, stream)
        self.solve_device(d, stream=stream)

    def goal_device(self, d: DeviceBatch, stream: int = 0):
        """goalPlanningWithGridBasedPlanner for every agent (traj_planner.cpp:545-550): d.goal_new, d.goal_status from
        the previous current_goal_point d.goal, the next waypoint and the planes of assemble_device"""
        self.qp.goal_batch(d.n, d.goal, d.waypoint, d.sfc, d.obs_offsets, d.normals, d.rhs, d.goal_new, d.goal_status,
                           stream=stream)

    def plan_device(self, d: DeviceBatch, generator: int = capi.GEN_CLSC, stream: int = 0):
        """One replan in the reference's stage order (TrajPlanner::planImpl, traj_planner.cpp:117-139):
        LSC construction with the previous goal, goal LP, then the QP with the new goal.  Agents whose goal LP is
        infeasible keep their previous goal and are reported through d.goal_status (the reference throws QPFAILED and
        keeps initial_traj for them); launches must be stream-ordered by the caller (same stream)."""
        import torch
        self.assemble_fused_device(d, generator, stream, prune=False)     # the goal LP needs every last-point plane
        self.goal_device(d, stream)
        ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
        with torch.cuda.stream(ext):
            ok = (d.goal_status == 0).unsqueeze(1)
            d.goal.copy_(torch.where(ok, d.goal_new, d.goal))
        self.solve_device(d, stream=stream)

    # ------------------------------------------------------------------ host-buffer path (what the reference calls)
    def host_buffers(self, batch: Batch) -> dict:
        """pinned copies of a batch's inputs + pinned outputs"""
        N, cfg = batch.n_agents, batch.cfg
        b = {}
        for name, arr, dt in (("state", batch.state, np.float32), ("goal", batch.goal, np.float32),
                              ("limits", batch.limits, np.float64), ("own_traj", batch.own_traj, np.float32),
                              ("agent_meta", batch.agent_meta, np.float64), ("next_waypoint", batch.next_waypoint, np.float32),
                              ("obs_offsets", batch.obs_offsets, np.int32), ("obs_index", batch.obs_index, np.int32)):
            p = capi.pinned(arr.shape, dt)
            p[...] = arr
            b[name] = p
        b["sfc"] = None
        if batch.sfc is not None:
            b["sfc"] = capi.pinned(batch.sfc.shape, np.float32)
            b["sfc"][...] = batch.sfc
        b["ctrl"] = capi.pinned((N, cfg.dim * cfg.M * 6), np.float64)
        b["cost"] = capi.pinned((N,), np.float64)
        b["status"] = capi.pinned((N,), np.int32)
        b["iters"] = capi.pinned((N,), np.int32)
        return b

    def replan_host_buffers(self, b: dict, n: int, generator: int = capi.GEN_LSC):
        self.qp.replan_host(generator, n, b["state"], b["goal"], b["limits"], b["sfc"], b["own_traj"], b["agent_meta"],
                            b["obs_offsets"], b["obs_index"], b["ctrl"], b["cost"], b["status"], b["iters"],
                            next_waypoint=b["next_waypoint"] if self.cfg.comm_range > 0 else None)

    def replan_host(self, batch: Batch, generator: int = capi.GEN_LSC) -> dict:
        b = self.host_buffers(batch)
        self.replan_host_buffers(b, batch.n_agents, generator)
        return {k: np.array(b[k]) for k in ("ctrl", "cost", "status", "iters")}

    @staticmethod
    def h2d_bytes(b: dict) -> int:
        return int(sum(b[k].nbytes for k in ("state", "goal", "limits", "own_traj", "agent_meta", "obs_offsets", "obs_index"))
                   + (b["sfc"].nbytes if b["sfc"] is not None else 0))

    @staticmethod
    def d2h_bytes(b: dict) -> int:
        return int(sum(b[k].nbytes for k in ("ctrl", "cost", "status", "iters")))
