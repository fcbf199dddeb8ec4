"""Closed-loop batched simulation (BASELINE config 5): N agents x T replans, agents sharded over ranks.

Per step, for the agents a rank owns (what MultiSyncSimulator::run does serially,
src/multi_sync_simulator.cpp:81-129):
    broadcastMsgs  -> neighbour lists from the all-gathered positions (K nearest within comm range)
    constructLSC   -> lscqp_assemble_lsc_fused: neighbours read in place, provably inactive pairs dropped (device)
    trajOptimization -> lscqp_solve_batch, started from the shifted previous solution (device)
    failsafe       -> agents whose QP failed keep initial_traj (src/traj_planner.cpp:767-797)
    doStep + exchange -> lscqp_step_exchange: failsafe, float trajectory, state at t = dt, shifted trajectory, stored
                      straight into every rank's exchange block over NVLink (CUDA IPC peer memory) + a sequence flag;
                      lscqp_exchange_begin at the start of the next step waits for all flags and copies the rows out.
                      The whole step is device-side and is captured in one CUDA graph.  exchange="nccl" keeps the
                      library all-gather (torch.distributed) as the comparison / fallback path.
Neighbour search runs in the library too (lscqp_select_neighbours: radix select in shared memory, one CTA per agent)."""
from __future__ import annotations

import numpy as np

from . import capi
from .sharding import all_agree, allgather_handles, allgather_rows, shard_range
from .workloads import Batch


class ClosedLoopSim:
    def __init__(self, batch: Batch, device: int = 0, rank: int = 0, world: int = 1, K: int = 40,
                 comm_range: float = 0.0, generator: int = capi.GEN_LSC, use_graph: bool = False,
                 goal_mode: str = "static", exchange: str = "p2p", world_boxes=None, world_resolution: float = 0.1):
        import torch
        from .planner import BatchPlanner
        self.torch = torch
        self.cfg = batch.cfg
        self.N = batch.n_agents
        self.K = min(K, self.N - 1, self.cfg.max_obs)
        self.comm_range = comm_range
        self.generator = generator
        self.rank, self.world = rank, world
        self.lo, self.hi = shard_range(self.N, rank, world)
        self.n_local = self.hi - self.lo
        self.planner = BatchPlanner(batch.cfg, device)
        dev = torch.device("cuda", device)
        self.dev = dev
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        M = self.cfg.M
        # global (replicated) per-agent data
        self.state = t(batch.state)                      # [N,9]
        self.goal = t(batch.goal)                        # [N,3] current_goal_point
        self.desired_goal = self.goal.clone()            # [N,3] desired_goal_point
        assert goal_mode in ("static", "righthand")
        self.goal_mode = goal_mode
        self._seq = torch.zeros((), dtype=torch.int32, device=dev)      # planner_seq, on the device (graph-safe)
        self.agent_meta = t(batch.agent_meta)            # [N,2]
        self.limits = t(batch.limits)
        # first replan: constant-velocity trajectories from the current state (traj_planner.cpp:276-279, 400-401)
        self.traj = self._const_vel(self.state)          # [N,M,6,3] initial_traj / obs_pred_trajs of this step
        # local scratch
        n, sk = max(self.n_local, 1), max(self.n_local * self.K, 1)
        self.obs_offsets = torch.zeros((self.n_local + 1,), dtype=torch.int32, device=dev)     # ragged CSR, rebuilt every step
        self.obs_index = torch.zeros((sk,), dtype=torch.int32, device=dev)
        self.overflow = torch.zeros((n,), dtype=torch.int32, device=dev)   # in-range count where it exceeds K (else 0)
        self._overflowed = torch.zeros((), dtype=torch.int64, device=dev)
        self.normals = torch.empty((sk, M, 3), dtype=torch.float64, device=dev)
        self.rhs = torch.empty((sk, M, 6), dtype=torch.float64, device=dev)
        self.ctrl = torch.empty((n, self.cfg.dim * M * 6), dtype=torch.float64, device=dev)
        self.cost = torch.empty((n,), dtype=torch.float64, device=dev)
        self.status = torch.empty((n,), dtype=torch.int32, device=dev)
        self.iters = torch.empty((n,), dtype=torch.int32, device=dev)
        self.traj_out = torch.empty((n, M, 6, 3), dtype=torch.float32, device=dev)
        self.state_out = torch.empty((n, 9), dtype=torch.float32, device=dev)
        self.shifted = torch.empty((n, M, 6, 3), dtype=torch.float32, device=dev)
        self._failed = torch.zeros((), dtype=torch.int64, device=dev)
        self.steps = 0
        self.use_graph = use_graph
        self._graph = None
        self.prune = bool(int(self.cfg.presolve) & 1)
        # Safe Flight Corridors against the static map (world_use_octomap; TrajPlanner::generateSFC, traj_planner.cpp:738-753)
        self.use_sfc = bool(self.cfg.use_sfc)
        self.sfc = None
        if self.use_sfc:
            assert world_boxes is not None, "cfg.use_sfc needs the world's boxes (a world CSV)"
            self.planner.qp.map_set(world_boxes, world_resolution, 1.0)
            wb = torch.tensor(list(self.cfg.world_min) + list(self.cfg.world_max), dtype=torch.float32, device=dev)
            self.sfc = wb.repeat(n, M, 1).contiguous()                  # (an agent without a valid start box keeps the world box)
            self.sfc_status = torch.zeros((n,), dtype=torch.int32, device=dev)
            self._sfc_invalid = torch.zeros((), dtype=torch.int64, device=dev)
        assert exchange in ("p2p", "nccl")
        self.exchange = exchange
        if exchange == "p2p":
            self._connect_peers()

    def _connect_peers(self):
        """every rank allocates its exchange block and maps the others' through CUDA IPC; the 64-byte handles travel
        through torch.distributed (plumbing).  If any rank cannot map a peer, all ranks fall back to the NCCL all-gather."""
        qp = self.planner.qp
        ok = True
        try:
            handle = qp.exchange_create(self.N, self.world, self.rank)
        except capi.LscqpError:
            ok, handle = False, bytes(64)
        if self.world > 1:
            handles, ok = allgather_handles(handle, ok, device=self.dev)
            if ok:
                try:
                    qp.exchange_connect(handles)
                except capi.LscqpError:
                    ok = False
            ok = all_agree(ok, device=self.dev)
        if not ok:
            self.exchange = "nccl"

    @property
    def failed_total(self) -> int:
        """QPs that did not converge so far (reads the device counter: synchronises)"""
        if self.exchange == "p2p":
            return self.planner.qp.exchange_status(self.torch.cuda.current_stream().cuda_stream)[2]
        return int(self._failed.item())

    @property
    def exchange_timeouts(self) -> int:
        return self.planner.qp.exchange_status(self.torch.cuda.current_stream().cuda_stream)[1] if self.exchange == "p2p" else 0

    @property
    def overflowed_total(self) -> int:
        """(agent, replan) pairs whose in-range neighbours exceeded the capacity K (reads the device counter)"""
        return int(self._overflowed.item())

    def _const_vel(self, state):
        torch = self.torch
        M, dt = self.cfg.M, self.cfg.dt
        tt = (torch.arange(M * 6, device=self.dev, dtype=torch.float32) * np.float32(dt / 5.0)).view(1, M, 6, 1)
        # Trajectory::planConstVelTraj (src/trajectory.cpp:77-89): time advances by dt/n per control point
        return (state[:, None, None, 0:3] + state[:, None, None, 3:6] * tt).contiguous()

    def neighbours(self, stream: int = 0):
        """the reference's obstacle set of every local agent (src/multi_sync_simulator.cpp:319-328: the agents within the
        L-inf communication range, all others when the range is <= 0) as a ragged CSR list, ids ascending; an agent with
        more than K in range keeps its K nearest and is counted in `overflowed_total` (lscqp_select_neighbours)"""
        self.planner.qp.select_neighbours(self.N, self.lo, self.n_local, self.K, self.comm_range, self.state,
                                          self.obs_offsets, self.obs_index, self.overflow, stream)
        return self.obs_index

    def _step_impl(self, stream: int):
        """one replan + one simulation step of length dt for the local shard, then the exchange; every launch goes to
        `stream` (the current torch stream), nothing returns to the host"""
        torch = self.torch
        qp = self.planner.qp
        n, lo, hi = self.n_local, self.lo, self.hi
        M_ = self.cfg.M
        p2p = self.exchange == "p2p"
        if p2p:
            qp.exchange_begin(self.traj, self.state, stream)           # rows every rank published at the end of the last step
        # goalPlanning (src/traj_planner.cpp:433-477): static goal, or the right-hand rule -- an agent that is slower than
        # deadlock/velocity_threshold (0.1) after deadlock/seq_threshold (5) replans and still more than 0.2 m from its
        # goal (isDeadlock, :904-923) aims at position + (desired - position) x e_z instead
        if self.goal_mode == "righthand":
            self._seq += 1                                             # (planner_seq: only the deadlock rule reads it)
            pos, vel = self.state[:, 0:3], self.state[:, 3:6]
            to_goal = self.desired_goal - pos
            dead = (self._seq > 5) & (vel.norm(dim=1) < 0.1) & (to_goal.norm(dim=1) > 0.2)
            side = torch.stack([to_goal[:, 1], -to_goal[:, 0], torch.zeros_like(to_goal[:, 2])], dim=1)
            self.goal.copy_(torch.where(dead[:, None], pos + side, self.desired_goal))
        assert n > 0, "every rank needs at least one agent"
        obs_index = self.neighbours(stream)
        own = self.traj[lo:hi]                                         # (contiguous row blocks: views, no copies)
        st, goal, lim, meta = self.state[lo:hi], self.goal[lo:hi], self.limits[lo:hi], self.agent_meta[lo:hi]
        qp.assemble_lsc_fused(self.generator, self.prune, n, own, meta, goal, st, lim, self.obs_offsets, obs_index,
                              self.traj, self.agent_meta, self.goal, self.state, self.normals, self.rhs, stream)
        if self.use_sfc:
            # constructSFC after constructLSC (planImpl, traj_planner.cpp:117-139): the first replan grows one box around the
            # start cell for all segments (:739-741), later ones shift and grow the last box from the end of initial_traj
            if self.steps == 0:
                qp.sfc_batch(capi.SFC_INIT, n, st[:, 0:3].contiguous(), None, None, lim, self.sfc, self.sfc_status, stream)
                self._sfc_invalid += (self.sfc_status == 0).sum()
            else:
                qp.sfc_batch(capi.SFC_FROM_POINT, n, own[:, M_ - 1, 5].contiguous(), goal, None, lim, self.sfc, self.sfc_status, stream)
        qp.solve_batch(n, st, goal, lim, self.sfc, self.obs_offsets, self.normals, self.rhs, self.ctrl, self.cost,
                       self.status, self.iters, stream=stream, initial_traj=own)
        self._overflowed += torch.count_nonzero(self.overflow[:n])    # (one reduction launch; overflow holds 0 or the count)
        if p2p:
            # failsafe (keep initial_traj where the QP did not converge, traj_planner.cpp:795-797) + doStep + shift +
            # publish to every rank, one kernel
            qp.step_exchange(lo, n, self.ctrl, self.status, own, self.cfg.dt, self.traj_out, stream)
            return
        bad = self.status != 0
        fallback = own.permute(0, 3, 1, 2)[:, :self.cfg.dim].reshape(n, -1).to(torch.float64)
        torch.where(bad[:, None], fallback, self.ctrl, out=self.ctrl)
        self._failed += bad.sum()
        qp.step_batch(n, self.ctrl, self.cfg.dt, self.traj_out, self.state_out, self.shifted, stream)
        if self.world == 1:
            self.traj.copy_(self.shifted[:n]); self.state.copy_(self.state_out[:n])

    def _exchange(self):
        """exchange="nccl": every rank ends the step with all trajectories and states through two library all-gathers
        (eager: NCCL collectives inside a captured graph hung intermittently at 2 ranks, so only the local compute is
        captured on this path; the p2p path has no such limit)"""
        n = self.n_local
        self.traj.copy_(allgather_rows(self.shifted[:n], self.N))
        self.state.copy_(allgather_rows(self.state_out[:n], self.N))

    def step(self):
        """One closed-loop step.  With use_graph the step (library launches and the few torch element-wise ops) is
        captured once into a CUDA graph after two eager steps and replayed afterwards: at a few hundred agents per GPU
        the step is otherwise bound by the host-side launch work.  On the p2p path the exchange is part of the graph."""
        torch = self.torch
        if self.use_graph and self._graph is not None:
            self._graph.replay()
        elif self.use_graph and self.steps >= 2:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_impl(torch.cuda.current_stream().cuda_stream)
            self._graph = g                  # (capture does not execute: replay performs this step)
            g.replay()
        else:
            self._step_impl(torch.cuda.current_stream().cuda_stream)
        if self.world > 1 and self.exchange == "nccl":
            self._exchange()
        self.steps += 1

    def sync_state(self):
        """make self.traj / self.state hold the rows of the last step (on the p2p path they are copied out of the inbox
        by the next step's first launch; call this before reading them on the host)"""
        if self.exchange == "p2p":
            self.planner.qp.exchange_begin(self.traj, self.state, self.torch.cuda.current_stream().cuda_stream)
        self.torch.cuda.synchronize()

    def min_separation_ratio(self, state=None) -> float:
        """min over pairs of (downwash-scaled distance) / (r_i + r_j) at the current positions (>= 1 is safe), or at the
        positions of a state snapshot"""
        torch = self.torch
        if state is None:
            self.sync_state()
            state = self.state
        pos = state[:, 0:3].to(torch.float64).clone()
        r = self.agent_meta[:, 0]; dw = self.agent_meta[:, 1]
        pos[:, 2] = pos[:, 2] / dw
        d = torch.cdist(pos, pos) / (r[:, None] + r[None, :])
        d.fill_diagonal_(float("inf"))
        return float(d.min())

    def max_goal_distance(self) -> float:
        self.sync_state()
        return float((self.state[:, 0:3] - self.desired_goal).norm(dim=1).max())
