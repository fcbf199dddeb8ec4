"""ctypes binding of liblscqp.so (include/lscqp.h).

This is plumbing only: it loads the in-tree CUDA library and passes pointers through.  There is
no CPU fallback -- if the library is missing or no CUDA device exists the calls raise.
Device entry points take torch CUDA tensors (their data_ptr()), host entry points take numpy
arrays (ideally backed by pinned memory, see `pinned`).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblscqp.so")

MODE_DLSC, MODE_LSC, MODE_BVC, MODE_RECIPROCALRSFC = 0, 1, 2, 4
GEN_LSC, GEN_CLSC, GEN_BVC, GEN_RSFC = 0, 1, 2, 3
SFC_INIT, SFC_FROM_POINT, SFC_FROM_HULL = 0, 1, 2
STATUS_OK, STATUS_MAX_ITER, STATUS_INFEASIBLE, STATUS_NUMERICAL, STATUS_CAPACITY = 0, 1, 2, 3, 4

EXPORTS = ["lscqp_version", "lscqp_last_error", "lscqp_create", "lscqp_destroy", "lscqp_dual_stride",
           "lscqp_max_obs_padded", "lscqp_assemble_lsc_batch", "lscqp_solve_batch", "lscqp_solve_host",
           "lscqp_replan_host", "lscqp_gather_obstacles", "lscqp_step_batch"]


class LscqpConfig(C.Structure):
    _fields_ = [("M", C.c_int), ("n", C.c_int), ("phi", C.c_int), ("dim", C.c_int),
                ("dt", C.c_double), ("w_control", C.c_double), ("w_terminal", C.c_double),
                ("planner_mode", C.c_int), ("use_sfc", C.c_int), ("comm_range", C.c_double),
                ("world_min", C.c_double * 3), ("world_max", C.c_double * 3), ("z_2d", C.c_double),
                ("max_obs", C.c_int), ("max_agents", C.c_int), ("max_iter", C.c_int), ("tol", C.c_double),
                ("presolve", C.c_int)]


class LscqpError(RuntimeError):
    pass


_lib = None


def load():
    """Load liblscqp.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LscqpError(f"{LIB_PATH} missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                             "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.lscqp_version.restype = C.c_char_p
        _lib.lscqp_last_error.restype = C.c_char_p
        _lib.lscqp_launch_count.restype = C.c_ulonglong
        _lib.lscqp_launch_count.argtypes = [C.c_void_p]
        _lib.lscqp_exchange_local_base.restype = C.c_void_p
        _lib.lscqp_exchange_local_base.argtypes = [C.c_void_p]
        for name in ("lscqp_solve_batch", "lscqp_assemble_lsc_batch", "lscqp_solve_host", "lscqp_replan_host",
                     "lscqp_gather_obstacles", "lscqp_step_batch", "lscqp_create", "lscqp_destroy", "lscqp_goal_batch",
                     "lscqp_goal_host", "lscqp_measure_fp64_peak", "lscqp_select_neighbours", "lscqp_assemble_lsc_fused", "lscqp_validate_batch",
                     "lscqp_last_instances", "lscqp_exchange_create", "lscqp_exchange_connect", "lscqp_exchange_begin",
                     "lscqp_step_exchange", "lscqp_exchange_status", "lscqp_exchange_destroy", "lscqp_exchange_connect_ptrs",
                     "lscqp_map_set", "lscqp_map_get", "lscqp_sfc_batch", "lscqp_sfc_host", "lscqp_set_obstacle_sizes"):
            getattr(_lib, name).restype = C.c_int
    return _lib


def make_config(cfg, max_agents: int = 0) -> LscqpConfig:
    """cfg: any object with the PlannerConfig fields (lsc_dr_planner_b200.workloads.PlannerConfig)."""
    return LscqpConfig(cfg.M, cfg.n, cfg.phi, cfg.dim, cfg.dt, cfg.w_control, cfg.w_terminal, cfg.planner_mode,
                       int(cfg.use_sfc), cfg.comm_range, (C.c_double * 3)(*cfg.world_min),
                       (C.c_double * 3)(*cfg.world_max), cfg.z_2d, cfg.max_obs, max_agents,
                       getattr(cfg, "max_iter", 0), getattr(cfg, "tol", 0.0), int(getattr(cfg, "presolve", True)))


def _dp(t):
    """device pointer of a torch tensor (or None)"""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "device entry points need contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def _hp(a, dtype):
    if a is None:
        return C.c_void_p(0)
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags["C_CONTIGUOUS"], (type(a), getattr(a, "dtype", None), dtype)
    return C.c_void_p(a.ctypes.data)


def pinned(shape, dtype):
    """numpy array backed by pinned host memory (through torch's allocator)."""
    import torch
    tdt = {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32}[np.dtype(dtype).type]
    t = torch.empty(shape, dtype=tdt, pin_memory=True)
    a = t.numpy()
    _PIN_KEEP.append(t)
    return a


_PIN_KEEP: list = []


class LscQp:
    """One handle = one (device, host thread).  Mirrors include/lscqp.h one to one."""

    def __init__(self, cfg, device: int = 0, max_agents: int = 0):
        self.lib = load()
        self.cfg = cfg
        self.ccfg = make_config(cfg, max_agents)
        self.h = C.c_void_p()
        rc = self.lib.lscqp_create(C.byref(self.ccfg), device, C.byref(self.h))
        self._check(rc)
        self.device = device
        self.dual_stride = self.lib.lscqp_dual_stride(self.h)
        self.kmax = self.lib.lscqp_max_obs_padded(self.h)
        self.nv = cfg.dim * cfg.M * 6

    def _check(self, rc):
        if rc != 0:
            raise LscqpError(f"lscqp error {rc}: {self.lib.lscqp_last_error().decode()}")

    def close(self):
        if self.h:
            self.lib.lscqp_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.lscqp_launch_count(self.h))

    def last_instances(self, n, stream=0) -> np.ndarray:
        """which pass solved every agent in the last two-pass solve_batch: 0 = the first pass (dual active-set kernel, or the
        light interior-point instance when that is the first pass), otherwise the full-capacity interior-point instance
        (after the active-set pass the value is its reason for deferring, das_kernel.cuh)"""
        out = np.zeros(n, np.int32)
        self._check(self.lib.lscqp_last_instances(self.h, n, _hp(out, np.int32), C.c_void_p(stream)))
        return out

    def measure_fp64_peak(self) -> float:
        """sustained FP64 FMA rate of the device in GFLOP/s (register-resident DFMA microbenchmark)"""
        out = C.c_double()
        self._check(self.lib.lscqp_measure_fp64_peak(self.h, C.byref(out)))
        return out.value

    # ------------------------------------------------------------------ device entry points
    def solve_batch(self, n, state, goal, limits, sfc, obs_offsets, normals, rhs, ctrl, cost, status,
                    iters=None, kkt=None, dual=None, stream=0, initial_traj=None, next_waypoint=None):
        self._check(self.lib.lscqp_solve_batch(self.h, n, _dp(state), _dp(goal), _dp(limits), _dp(sfc), _dp(next_waypoint),
                                               _dp(obs_offsets),
                                               _dp(normals), _dp(rhs), _dp(initial_traj), _dp(ctrl), _dp(cost), _dp(status), _dp(iters),
                                               _dp(kkt), _dp(dual), C.c_void_p(stream)))

    def assemble_lsc_batch(self, generator, n, own_traj, agent_meta, agent_goal, obs_offsets, obs_traj, obs_meta,
                           obs_goal, obs_position, normals, rhs, stream=0):
        self._check(self.lib.lscqp_assemble_lsc_batch(self.h, generator, n, _dp(own_traj), _dp(agent_meta), _dp(agent_goal),
                                                      _dp(obs_offsets), _dp(obs_traj), _dp(obs_meta), _dp(obs_goal),
                                                      _dp(obs_position), _dp(normals), _dp(rhs), C.c_void_p(stream)))

    def set_obstacle_sizes(self, obs_size):
        """predicted obstacle sizes [sumK][M][6] (device tensor) for GEN_RSFC, or None = the obstacles' radii"""
        self._obs_size_keep = obs_size
        self._check(self.lib.lscqp_set_obstacle_sizes(self.h, _dp(obs_size)))

    def assemble_lsc_fused(self, generator, prune, n, own_traj, agent_meta, agent_goal, state, limits, obs_offsets, obs_index,
                           all_traj, all_meta, all_goal, all_state, normals, rhs, stream=0):
        """gather-free assembly for obstacles that are agents of the same population, optionally pruned (exact)"""
        self._check(self.lib.lscqp_assemble_lsc_fused(self.h, generator, int(prune), n, _dp(own_traj), _dp(agent_meta),
                                                      _dp(agent_goal), _dp(state), _dp(limits), _dp(obs_offsets),
                                                      _dp(obs_index), _dp(all_traj), _dp(all_meta), _dp(all_goal),
                                                      _dp(all_state), _dp(normals), _dp(rhs), C.c_void_p(stream)))

    def gather_obstacles(self, n_obs, obs_index, own_traj, agent_meta, agent_goal, state, obs_traj, obs_meta, obs_goal,
                         obs_position, stream=0):
        self._check(self.lib.lscqp_gather_obstacles(self.h, n_obs, _dp(obs_index), _dp(own_traj), _dp(agent_meta),
                                                    _dp(agent_goal), _dp(state), _dp(obs_traj), _dp(obs_meta),
                                                    _dp(obs_goal), _dp(obs_position), C.c_void_p(stream)))

    def validate_batch(self, n, traj, state_at_step, limits, sfc, valid_out, stream=0):
        """TrajPlanner::isSolValid for every agent (traj_planner.cpp:990-1045)"""
        self._check(self.lib.lscqp_validate_batch(self.h, n, _dp(traj), _dp(state_at_step), _dp(limits), _dp(sfc),
                                                  _dp(valid_out), C.c_void_p(stream)))

    def select_neighbours(self, n_total, lo, n_local, K, comm_range, state, obs_offsets_out, obs_index_out,
                          overflow_out=None, stream=0):
        """broadcastMsgs on the device: ragged CSR lists of the agents within the Chebyshev communication range of the agents
        [lo, lo + n_local) (all others when comm_range <= 0); more than K in range -> K nearest + overflow_out = count"""
        self._check(self.lib.lscqp_select_neighbours(self.h, n_total, lo, n_local, K, C.c_double(comm_range), _dp(state),
                                                     _dp(obs_offsets_out), _dp(obs_index_out), _dp(overflow_out),
                                                     C.c_void_p(stream)))

    def step_batch(self, n, ctrl, step, traj_out, state_out=None, shifted_out=None, stream=0):
        self._check(self.lib.lscqp_step_batch(self.h, n, _dp(ctrl), C.c_double(step), _dp(traj_out), _dp(state_out),
                                              _dp(shifted_out), C.c_void_p(stream)))

    # ------------------------------------------------------------------ static map + Safe Flight Corridors
    def map_set(self, boxes, resolution: float = 0.1, max_dist: float = 1.0):
        """world boxes [n][6] (centre xyz, size xyz: the rows of a world CSV) -> occupancy grid + nearest-obstacle field
        on the device (MapManager::updateOctreeFromCSV + DynamicEDTOctomap(maxdist))"""
        b = np.ascontiguousarray(np.asarray(boxes, np.float64).reshape(-1, 6))
        self._check(self.lib.lscqp_map_set(self.h, _hp(b, np.float64) if b.size else C.c_void_p(0), b.shape[0],
                                           C.c_double(resolution), C.c_double(max_dist)))

    def map_get(self):
        """(occupancy [nx, ny, nz] uint8, closest [nx, ny, nz] packed int32) copied back from the device"""
        n3 = (C.c_int * 3)()
        self._check(self.lib.lscqp_map_get(self.h, n3, C.c_void_p(0), C.c_void_p(0)))
        occ = np.zeros(tuple(n3), np.uint8); cl = np.zeros(tuple(n3), np.int32)
        self._check(self.lib.lscqp_map_get(self.h, n3, C.c_void_p(occ.ctypes.data), C.c_void_p(cl.ctypes.data)))
        return occ, cl

    def sfc_batch(self, mode, n, point, goal, next_waypoint, limits, sfc, status, stream=0):
        """TrajPlanner::generateSFC for every agent: SFC_INIT / SFC_FROM_POINT / SFC_FROM_HULL on sfc [n][M][6] (device)"""
        self._check(self.lib.lscqp_sfc_batch(self.h, mode, n, _dp(point), _dp(goal), _dp(next_waypoint), _dp(limits), _dp(sfc),
                                             _dp(status), C.c_void_p(stream)))

    # ------------------------------------------------------------------ peer exchange (sharded closed loop)
    def exchange_create(self, n_total, world, rank) -> bytes:
        """allocate this rank's exchange block; returns its 64-byte CUDA IPC handle (to be all-gathered by the caller)"""
        buf = (C.c_ubyte * 64)()
        self._check(self.lib.lscqp_exchange_create(self.h, n_total, world, rank, buf))
        return bytes(buf)

    def exchange_connect(self, all_handles: bytes):
        """all_handles: world x 64 bytes in rank order"""
        self._check(self.lib.lscqp_exchange_connect(self.h, C.c_char_p(all_handles)))

    def exchange_begin(self, traj, state, stream=0):
        self._check(self.lib.lscqp_exchange_begin(self.h, _dp(traj), _dp(state), C.c_void_p(stream)))

    def step_exchange(self, lo, n_local, ctrl, status, fallback_traj, step, traj_out=None, stream=0):
        self._check(self.lib.lscqp_step_exchange(self.h, lo, n_local, _dp(ctrl), _dp(status), _dp(fallback_traj),
                                                 C.c_double(step), _dp(traj_out), C.c_void_p(stream)))

    def exchange_status(self, stream=0) -> tuple[int, int, int]:
        """(steps published, wait time-outs, failsafe uses) of this rank; synchronises the stream"""
        out = (C.c_ulonglong * 3)()
        self._check(self.lib.lscqp_exchange_status(self.h, out, C.c_void_p(stream)))
        return int(out[0]), int(out[1]), int(out[2])

    def goal_batch(self, n, goal, next_waypoint, sfc, obs_offsets, normals, rhs, goal_out, status, t_out=None, stream=0):
        """batched GoalOptimizer::solve (goal_optimizer.cpp:7-165) on device tensors"""
        self._check(self.lib.lscqp_goal_batch(self.h, n, _dp(goal), _dp(next_waypoint), _dp(sfc), _dp(obs_offsets),
                                              _dp(normals), _dp(rhs), _dp(goal_out), _dp(t_out), _dp(status),
                                              C.c_void_p(stream)))

    # ------------------------------------------------------------------ host entry points
    def goal_host(self, n, goal, next_waypoint, sfc, obs_offsets, normals, rhs, goal_out, status, t_out=None):
        self._check(self.lib.lscqp_goal_host(self.h, n, _hp(goal, np.float32), _hp(next_waypoint, np.float32),
                                             _hp(sfc, np.float32), _hp(obs_offsets, np.int32), _hp(normals, np.float64),
                                             _hp(rhs, np.float64), _hp(goal_out, np.float32), _hp(t_out, np.float64),
                                             _hp(status, np.int32)))

    def solve_host(self, n, state, goal, limits, sfc, obs_offsets, normals, rhs, ctrl, cost, status, iters=None,
                   kkt=None, dual=None, initial_traj=None, next_waypoint=None):
        self._check(self.lib.lscqp_solve_host(self.h, n, _hp(state, np.float32), _hp(goal, np.float32),
                                              _hp(limits, np.float64), _hp(sfc, np.float32),
                                              _hp(next_waypoint, np.float32), _hp(obs_offsets, np.int32),
                                              _hp(normals, np.float64), _hp(rhs, np.float64), _hp(initial_traj, np.float32),
                                              _hp(ctrl, np.float64),
                                              _hp(cost, np.float64), _hp(status, np.int32), _hp(iters, np.int32),
                                              _hp(kkt, np.float64), _hp(dual, np.float64)))

    def replan_host(self, generator, n, state, goal, limits, sfc, own_traj, agent_meta, obs_offsets, obs_index, ctrl,
                    cost, status, iters=None, next_waypoint=None):
        self._check(self.lib.lscqp_replan_host(self.h, generator, n, _hp(state, np.float32), _hp(goal, np.float32),
                                               _hp(limits, np.float64), _hp(sfc, np.float32),
                                               _hp(next_waypoint, np.float32), _hp(own_traj, np.float32),
                                               _hp(agent_meta, np.float64), _hp(obs_offsets, np.int32),
                                               _hp(obs_index, np.int32), _hp(ctrl, np.float64), _hp(cost, np.float64),
                                               _hp(status, np.int32), _hp(iters, np.int32)))
