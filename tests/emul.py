"""ctypes face of tests/cuda_emul/_build/libemul.so: the kernel sources run on the CPU thread emulator.
TEST INFRASTRUCTURE ONLY -- never imported by the package."""
import ctypes as C
import os

import numpy as np

from lsc_dr_planner_b200 import capi

_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emul", "_build", "libemul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a, dt):
    if a is None:
        return C.c_void_p(0)
    assert a.dtype == dt and a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def solve_batch(cfg, n, state, goal, limits, sfc, off, normals, rhs, want_dual=False, initial_traj=None, next_waypoint=None):
    cc = capi.make_config(cfg)
    nv = cfg.dim * cfg.M * 6
    ctrl = np.zeros((n, nv)); cost = np.zeros(n); status = np.zeros(n, np.int32); iters = np.zeros(n, np.int32)
    kkt = np.zeros((n, 4))
    ds = lib().emul_dual_stride(cfg.M, cfg.dim, int(cfg.comm_range > 0))
    dual = np.zeros((n, ds)) if want_dual else None
    rc = lib().emul_solve_batch(C.byref(cc), n, _p(state, np.float32), _p(goal, np.float32), _p(limits, np.float64),
                                _p(sfc, np.float32), _p(next_waypoint, np.float32), _p(off, np.int32), _p(normals, np.float64), _p(rhs, np.float64),
                                _p(initial_traj, np.float32),
                                _p(ctrl, np.float64), _p(cost, np.float64), _p(status, np.int32), _p(iters, np.int32),
                                _p(kkt, np.float64), _p(dual, np.float64))
    assert rc == 0, rc
    return ctrl, cost, status, iters, kkt, dual


def assemble(cfg, generator, n, own_traj, agent_meta, agent_goal, off, obs_traj, obs_meta, obs_goal, obs_position, obs_size=None):
    cc = capi.make_config(cfg)
    lib().emul_set_obstacle_sizes(_p(obs_size, np.float64))
    sk = int(off[n])
    normals = np.zeros((sk, cfg.M, 3)); rhs = np.zeros((sk, cfg.M, 6))
    rc = lib().emul_assemble_lsc_batch(C.byref(cc), generator, n, _p(own_traj, np.float32), _p(agent_meta, np.float64),
                                       _p(agent_goal, np.float32), _p(off, np.int32), _p(obs_traj, np.float32),
                                       _p(obs_meta, np.float32), _p(obs_goal, np.float32), _p(obs_position, np.float32),
                                       _p(normals, np.float64), _p(rhs, np.float64))
    lib().emul_set_obstacle_sizes(C.c_void_p(0))
    assert rc == 0, rc
    return normals, rhs


def step(cfg, n, ctrl, t, status=None, fallback=None):
    cc = capi.make_config(cfg)
    traj = np.zeros((n, cfg.M, 6, 3), np.float32); state = np.zeros((n, 9), np.float32); shifted = np.zeros_like(traj)
    rc = lib().emul_step_batch(C.byref(cc), n, _p(ctrl, np.float64), C.c_double(t), _p(traj, np.float32),
                               _p(state, np.float32), _p(shifted, np.float32), _p(status, np.int32), _p(fallback, np.float32))
    assert rc == 0, rc
    return traj, state, shifted


class ExchangeSim:
    """`world` ranks of the peer exchange simulated in one process on the emulator (step_kernel publishing into every
    rank's block, exchange_begin_kernel copying the inbox out)"""

    def __init__(self, cfg, world, n_total):
        self.cfg, self.world, self.n = cfg, world, n_total
        lib().emul_exchange_bytes.restype = C.c_long
        nb = lib().emul_exchange_bytes(n_total, cfg.M)
        self.blocks = [np.zeros(nb // 8 + 1, np.uint64) for _ in range(world)]
        self.ptrs = (C.c_void_p * world)(*[b.ctypes.data for b in self.blocks])
        self.traj = np.zeros((world, n_total, cfg.M, 6, 3), np.float32)
        self.state = np.zeros((world, n_total, 9), np.float32)

    def step(self, ctrl, t, status=None, fallback=None, skip_rank=-1):
        cc = capi.make_config(self.cfg)
        rc = lib().emul_exchange_step(C.byref(cc), self.world, self.n, self.ptrs, _p(ctrl, np.float64), _p(status, np.int32),
                                      _p(fallback, np.float32), C.c_double(t), _p(self.traj, np.float32),
                                      _p(self.state, np.float32), skip_rank)
        assert rc == 0, rc

    def counters(self, rank):
        """(steps published, time-outs, failsafe uses)"""
        c = self.blocks[rank]
        return int(c[0]), int(c[2]), int(c[3])


def goal(cfg, n, goal_pt, waypoint, sfc, off, normals, rhs):
    cc = capi.make_config(cfg)
    out = np.zeros((n, 3), np.float32); t = np.zeros(n); status = np.zeros(n, np.int32)
    rc = lib().emul_goal_batch(C.byref(cc), n, _p(goal_pt, np.float32), _p(waypoint, np.float32), _p(sfc, np.float32),
                               _p(off, np.int32), _p(normals, np.float64), _p(rhs, np.float64), _p(out, np.float32),
                               _p(t, np.float64), _p(status, np.int32))
    assert rc == 0, rc
    return out, t, status


def select_neighbours(n_total, lo, n_local, K, comm_range, state):
    """-> (offsets [n_local+1], ids [sumK], overflow [n_local])"""
    off = np.zeros(n_local + 1, np.int32); idx = np.full(max(n_local * K, 1), -1, np.int32); over = np.zeros(n_local, np.int32)
    rc = lib().emul_select_neighbours(n_total, lo, n_local, K, C.c_double(comm_range), _p(state, np.float32), _p(off, np.int32),
                                      _p(idx, np.int32), _p(over, np.int32))
    assert rc == 0, rc
    return off, idx[:off[-1]], over


def assemble_fused(cfg, generator, prune, batch):
    """gather-free (optionally pruned) assembly of a whole workloads.Batch whose obstacles are its own agents"""
    cc = capi.make_config(cfg)
    n = batch.n_agents; sk = int(batch.obs_offsets[n])
    normals = np.full((sk, cfg.M, 3), np.nan); rhs = np.full((sk, cfg.M, 6), np.nan)
    meta = np.ascontiguousarray(batch.agent_meta, np.float64)
    rc = lib().emul_assemble_lsc_fused(C.byref(cc), generator, int(prune), n, _p(batch.own_traj, np.float32), _p(meta, np.float64),
                                       _p(batch.goal, np.float32), _p(batch.state, np.float32), _p(batch.limits, np.float64),
                                       _p(batch.obs_offsets, np.int32), _p(batch.obs_index, np.int32), _p(batch.own_traj, np.float32),
                                       _p(meta, np.float64), _p(batch.goal, np.float32), _p(batch.state, np.float32),
                                       _p(normals, np.float64), _p(rhs, np.float64))
    assert rc == 0, rc
    return normals, rhs


def validate(cfg, n, traj, state, limits, sfc):
    cc = capi.make_config(cfg)
    out = np.zeros(n, np.int32)
    rc = lib().emul_validate_batch(C.byref(cc), n, _p(traj, np.float32), _p(state, np.float32), _p(limits, np.float64),
                                   _p(sfc, np.float32), _p(out, np.int32))
    assert rc == 0, rc
    return out


class EmulMap:
    """the map kernels (occupancy_kernel, edt_closest_kernel) and sfc_kernel on the emulator"""

    def __init__(self, cfg, boxes, resolution=0.1, max_dist=1.0):
        self.cfg = cfg
        self.cc = capi.make_config(cfg)
        self.boxes = np.ascontiguousarray(np.asarray(boxes, np.float64).reshape(-1, 6))
        lib().emul_map_build.restype = C.c_void_p
        lib().emul_map_occ.restype = C.c_void_p
        lib().emul_map_closest.restype = C.c_void_p
        self.h = C.c_void_p(lib().emul_map_build(C.byref(self.cc), _p(self.boxes, np.float64), self.boxes.shape[0],
                                                 C.c_double(resolution), C.c_double(max_dist)))
        n3 = (C.c_int * 3)()
        self._occ_ptr = lib().emul_map_occ(self.h, n3)
        self.n = tuple(n3)

    def __del__(self):
        try:
            lib().emul_map_free(self.h)
        except Exception:
            pass

    def occupancy(self):
        return np.ctypeslib.as_array(C.cast(self._occ_ptr, C.POINTER(C.c_ubyte)), shape=self.n).copy()

    def closest_view(self):
        """packed nearest-occupied-cell table [nx, ny, nz] (writable view)"""
        return np.ctypeslib.as_array(C.cast(lib().emul_map_closest(self.h), C.POINTER(C.c_int)), shape=self.n)

    def edt_cells(self, cells):
        """run edt_closest_kernel for the CTAs holding the given linear cell indices"""
        c = np.ascontiguousarray(cells, np.int64)
        assert lib().emul_map_edt_cells(self.h, c.ctypes.data_as(C.POINTER(C.c_long)), len(c)) == 0

    def sfc(self, mode, point, goal, waypoint, limits, sfc):
        n = point.shape[0]
        status = np.full(n, -1, np.int32)
        rc = lib().emul_sfc_batch(self.h, mode, self.cfg.M, n, _p(point, np.float32), _p(goal, np.float32), _p(waypoint, np.float32),
                                  _p(limits, np.float64), _p(sfc, np.float32), _p(status, np.int32))
        assert rc == 0
        return status
