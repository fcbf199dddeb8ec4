"""oracle/oracle.py -- Python face of the CPU checker.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing under lsc_dr_planner_b200/ does.

What it wraps
  * oracle/lscqp_oracle.c  -- C restatement of the reference path
    (src/traj_optimizer.cpp:163-538, src/traj_planner.cpp:611-736, include/geometry.hpp,
    src/trajectory.cpp), built by oracle/Makefile into oracle/_build/liblscqp_oracle.so.
  * oracle/_ref/libopengjk_ref.so -- the reference's own openGJK (when built here).
  * HiGHS 1.12 (SciPy-bundled, private API scipy.optimize._highspy._core) standing in for
    IBM CPLEX 20.1, which the reference calls at src/traj_optimizer.cpp:66 and which is
    absent from this image.  QP-solution parity is therefore UNPINNED against CPLEX; it is
    certified through KKT residuals on the restated model (unique minimiser).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liblscqp_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libopengjk_ref.so")
INF = 1e30

MODE_DLSC, MODE_LSC, MODE_BVC, MODE_ORCA, MODE_RECIPROCALRSFC = 0, 1, 2, 3, 4
GEN_LSC, GEN_CLSC, GEN_BVC, GEN_RSFC = 0, 1, 2, 3


def build(force: bool = False) -> None:
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(
            os.path.join(_HERE, f)) for f in ("lscqp_oracle.c", "sfc_oracle.c", "lscqp_oracle.h")):
        subprocess.run(["make", "-C", _HERE, "_build/liblscqp_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/src/openGJK/openGJK.cpp") and (force or not os.path.exists(_REF)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.DEVNULL)


_PDIP = os.path.join(_HERE, "_build", "libpdip_cpu.so")


def _cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build_pdip(force: bool = False) -> None:
    """Compile oracle/pdip_cpu.c with -O3 -march=native -fopenmp for THIS machine (rebuilt when the CPU model changes:
    the prebuilt file travels to the GPU box, whose host CPU may differ)."""
    stamp = _PDIP + ".cpu"
    srcs = [os.path.join(_HERE, f) for f in ("pdip_cpu.c", "lscqp_oracle.c", "lscqp_oracle.h")]
    stale = (not os.path.exists(_PDIP) or not os.path.exists(stamp) or open(stamp).read() != _cpu_model()
             or os.path.getmtime(_PDIP) < max(os.path.getmtime(f) for f in srcs))
    if force or stale:
        if os.path.exists(_PDIP):
            os.unlink(_PDIP)
        subprocess.run(["make", "-C", _HERE, "_build/libpdip_cpu.so"], check=True, stdout=subprocess.DEVNULL)
        open(stamp, "w").write(_cpu_model())


_pdip = None


def pdip_lib():
    global _pdip
    if _pdip is None:
        build_pdip()
        _pdip = C.CDLL(_PDIP)
        _pdip.orc_pdip_solve.restype = C.c_int
        _pdip.orc_replan_batch_pdip.restype = C.c_double
    return _pdip


class _Cfg(C.Structure):
    _fields_ = [("M", C.c_int), ("n", C.c_int), ("phi", C.c_int), ("phi_n", C.c_int), ("dim", C.c_int),
                ("dt", C.c_double), ("w_control", C.c_double), ("w_terminal", C.c_double),
                ("planner_mode", C.c_int), ("use_sfc", C.c_int), ("comm_range", C.c_double),
                ("world_min", C.c_double * 3), ("world_max", C.c_double * 3), ("z_2d", C.c_double)]


class _Agent(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("velocity", C.c_float * 3), ("acceleration", C.c_float * 3),
                ("current_goal_point", C.c_float * 3), ("next_waypoint", C.c_float * 3),
                ("max_vel", C.c_double * 3), ("max_acc", C.c_double * 3),
                ("radius", C.c_double), ("nominal_velocity", C.c_double)]


@dataclass
class Config:
    M: int = 5
    n: int = 5
    phi: int = 3
    phi_n: int = 1
    dim: int = 3
    dt: float = 0.2
    w_control: float = 0.01
    w_terminal: float = 1.0
    planner_mode: int = MODE_LSC
    use_sfc: bool = False
    comm_range: float = 0.0
    world_min: tuple = (-10.0, -10.0, 0.0)
    world_max: tuple = (10.0, 10.0, 2.5)
    z_2d: float = 1.0

    def c(self) -> _Cfg:
        return _Cfg(self.M, self.n, self.phi, self.phi_n, self.dim, self.dt, self.w_control, self.w_terminal,
                    self.planner_mode, int(self.use_sfc), self.comm_range,
                    (C.c_double * 3)(*self.world_min), (C.c_double * 3)(*self.world_max), self.z_2d)


@dataclass
class Agent:
    position: np.ndarray
    velocity: np.ndarray
    acceleration: np.ndarray
    goal: np.ndarray
    next_waypoint: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    max_vel: tuple = (1.0, 1.0, 1.0)
    max_acc: tuple = (2.0, 2.0, 2.0)
    radius: float = 0.15
    nominal_velocity: float = 1.0
    downwash: float = 2.0

    def c(self) -> _Agent:
        f3 = lambda a: (C.c_float * 3)(*np.asarray(a, np.float32).tolist())
        return _Agent(f3(self.position), f3(self.velocity), f3(self.acceleration), f3(self.goal),
                      f3(self.next_waypoint), (C.c_double * 3)(*self.max_vel), (C.c_double * 3)(*self.max_acc),
                      self.radius, self.nominal_velocity)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_min_norm_hull.restype = C.c_double
        _lib.orc_terminal_segments.restype = C.c_int
        _lib.orc_qp_build.restype = C.c_int
        _lib.orc_build_aeq_base.restype = C.c_int
        _lib.orc_map_build.restype = C.c_void_p
        _lib.orc_map_occupancy.restype = C.c_void_p
        _lib.orc_map_closest.restype = C.c_void_p
        for f in ("orc_map_free", "orc_map_dims", "orc_map_occupancy", "orc_map_closest", "orc_is_obstacle_in_sfc", "orc_expand_sfc",
                  "orc_sfc_initialize", "orc_sfc_from_point", "orc_sfc_from_convex_hull"):
            getattr(_lib, f).argtypes = None
    return _lib


def ref_available() -> bool:
    return os.path.exists(_REF)


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF)
        _ref.ref_gjk_hull_origin.restype = C.c_double
    return _ref


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


# ---------------------------------------------------------------- constants
def bernstein_basis(n: int) -> np.ndarray:
    B = np.zeros((n + 1, n + 1))
    lib().orc_bernstein_basis(n, _p(B, C.c_double))
    return B


def qbase(n=5, phi=3, phi_n=1, dt=0.2) -> np.ndarray:
    Q = np.zeros((n + 1, n + 1))
    lib().orc_build_qbase(n, phi, phi_n, C.c_double(dt), _p(Q, C.c_double))
    return Q


def aeq_base(M, n=5, phi=3, dt=0.2) -> np.ndarray:
    A = np.zeros(((M - 2) * phi, M * (n + 1)))
    rc = lib().orc_build_aeq_base(M, n, phi, C.c_double(dt), _p(A, C.c_double))
    if rc != 0:
        raise ValueError("[TrajOptimizer] Currently, only n=5, phi=3 is available")
    return A


def terminal_segments(cfg: Config, ag: Agent) -> int:
    cc, ca = cfg.c(), ag.c()
    return lib().orc_terminal_segments(C.byref(cc), C.byref(ca))


# ---------------------------------------------------------------- QP model
@dataclass
class QP:
    P: np.ndarray       # objective x'Px + q'x + c0 (no 1/2)
    q: np.ndarray
    c0: float
    Aeq: np.ndarray
    beq: np.ndarray
    G: np.ndarray       # rlo <= Gx <= rhi
    rlo: np.ndarray
    rhi: np.ndarray
    lb: np.ndarray
    ub: np.ndarray


def qp_build(cfg: Config, ag: Agent, lsc_point, lsc_normal, lsc_d, sfc=None) -> QP:
    """Dense populatebyrow (src/traj_optimizer.cpp:216-514).  lsc_*: [K][M][n+1](,3)."""
    lsc_point, lsc_normal, lsc_d = _f32(lsc_point), _f32(lsc_normal), _f64(lsc_d)
    K = lsc_d.shape[0] if lsc_d.size else 0
    nv, ne, ni = C.c_int(), C.c_int(), C.c_int()
    cc, ca = cfg.c(), ag.c()
    lib().orc_qp_sizes(C.byref(cc), K, _p(lsc_normal, C.c_float) if K else None,
                       C.byref(nv), C.byref(ne), C.byref(ni))
    nv, ne, ni = nv.value, ne.value, ni.value
    P = np.zeros((nv, nv)); q = np.zeros(nv); c0 = C.c_double()
    A = np.zeros((ne, nv)); b = np.zeros(ne)
    G = np.zeros((ni, nv)); rlo = np.zeros(ni); rhi = np.zeros(ni)
    lb = np.zeros(nv); ub = np.zeros(nv)
    if sfc is None:
        sfc = np.zeros((cfg.M, 6), np.float32)
    sfc = _f32(sfc)
    rc = lib().orc_qp_build(C.byref(cc), C.byref(ca), K, _p(lsc_point, C.c_float), _p(lsc_normal, C.c_float),
                            _p(lsc_d, C.c_double), _p(sfc, C.c_float),
                            _p(P, C.c_double), _p(q, C.c_double), C.byref(c0), _p(A, C.c_double), _p(b, C.c_double),
                            _p(G, C.c_double), _p(rlo, C.c_double), _p(rhi, C.c_double),
                            _p(lb, C.c_double), _p(ub, C.c_double))
    if rc != 0:
        raise ValueError(f"orc_qp_build failed rc={rc}")
    return QP(P, q, c0.value, A, b, G, rlo, rhi, lb, ub)


# ---------------------------------------------------------------- geometry / assembly
def min_norm_hull(pts) -> tuple[np.ndarray, float]:
    pts = _f64(pts)
    v = np.zeros(3)
    d = lib().orc_min_norm_hull(_p(pts, C.c_double), pts.shape[0], _p(v, C.c_double))
    return v, d


def ref_gjk(pts) -> tuple[np.ndarray, float]:
    """The reference's own openGJK on (hull, origin) -- geometry.hpp:266-296."""
    pts = _f64(pts)
    v = np.zeros(3)
    d = ref().ref_gjk_hull_origin(_p(pts, C.c_double), pts.shape[0], _p(v, C.c_double))
    return v, d


def closest_points_segments(l1s, l1e, l2s, l2e):
    a, b, c, d = (_f32(x) for x in (l1s, l1e, l2s, l2e))
    cp1 = np.zeros(3, np.float32); cp2 = np.zeros(3, np.float32); dist = C.c_double()
    lib().orc_closest_points_segments(_p(a, C.c_float), _p(b, C.c_float), _p(c, C.c_float), _p(d, C.c_float),
                                      _p(cp1, C.c_float), _p(cp2, C.c_float), C.byref(dist))
    return cp1, cp2, dist.value


def obstacle_sizes(cfg: Config, obs_radius: float, obs_max_acc: float, uncertainty_horizon: float = 1.0, velocity_guard: float = 0.0):
    """obstacleSizePredictionWithConstAcc (src/traj_planner.cpp:321-358) for one obstacle: [M, n+1] predicted sizes"""
    out = np.zeros((cfg.M, cfg.n + 1))
    lib().orc_obstacle_sizes(cfg.M, cfg.n, C.c_double(cfg.dt), C.c_double(obs_radius), C.c_double(obs_max_acc),
                             C.c_double(uncertainty_horizon), C.c_double(velocity_guard), _p(out, C.c_double))
    return out


def generate_lsc(cfg: Config, generator: int, ag: Agent, own_traj, obs_traj, obs_radius, obs_downwash,
                 obs_goal=None, obs_position=None, obs_size=None):
    """generateLSC / generateCLSC / generateBVC / generateReciprocalRSFC (src/traj_planner.cpp:581-736).
    obs_size [K, M, n+1]: predicted obstacle sizes (GEN_RSFC; None: the obstacle radius).
    Returns (point[K,M,6,3] f32, normal[K,M,6,3] f32, d[K,M,6] f64)."""
    own_traj, obs_traj = _f32(own_traj), _f32(obs_traj)
    K = obs_traj.shape[0]
    N = cfg.n + 1
    obs_radius, obs_downwash = _f32(obs_radius), _f32(obs_downwash)
    obs_goal = _f32(np.zeros((K, 3)) if obs_goal is None else obs_goal)
    obs_position = _f32(obs_traj[:, 0, 0, :] if obs_position is None else obs_position)
    pt = np.zeros((K, cfg.M, N, 3), np.float32); nr = np.zeros((K, cfg.M, N, 3), np.float32)
    d = np.zeros((K, cfg.M, N))
    cc, ca = cfg.c(), ag.c()
    slot = C.c_void_p.in_dll(lib(), "orc_rsfc_sizes")
    sizes = None
    if obs_size is not None:
        sizes = np.ascontiguousarray(obs_size, np.float64).reshape(K, cfg.M, N)
        slot.value = sizes.ctypes.data
    try:
        lib().orc_generate_lsc(C.byref(cc), generator, C.byref(ca), C.c_double(ag.downwash), _p(own_traj, C.c_float), K,
                               _p(obs_traj, C.c_float), _p(obs_radius, C.c_float), _p(obs_downwash, C.c_float),
                               _p(obs_goal, C.c_float), _p(obs_position, C.c_float),
                               _p(pt, C.c_float), _p(nr, C.c_float), _p(d, C.c_double))
    finally:
        slot.value = None
    return pt, nr, d


def pack_planes(cfg: Config, lsc_point, lsc_normal, lsc_d):
    """Packed form the kernels consume: normal[K,M,3] f64 and rhs b = n.p + d [K,M,6] f64
    (the constant the reference's row n.(c - p) - d >= 0 carries, traj_optimizer.cpp:413-429).
    Requires the normal to be shared by the n+1 records of one (oi, m), as every generator does."""
    nr = np.asarray(lsc_normal, np.float64); pt = np.asarray(lsc_point, np.float64)
    assert np.all(nr == nr[:, :, :1, :]), "normals differ inside one (obstacle, segment)"
    normal = nr[:, :, 0, :].copy()
    if cfg.dim == 2:
        rhs = np.einsum("kmd,kmid->kmi", normal[..., :2], pt[..., :2]) + np.asarray(lsc_d)
    else:
        rhs = np.einsum("kmd,kmid->kmi", normal, pt) + np.asarray(lsc_d)
    return normal, rhs


# ---------------------------------------------------------------- closed-loop glue
def get_state_at(cfg: Config, traj, time: float) -> np.ndarray:
    traj = _f32(traj); out = np.zeros(9, np.float32)
    lib().orc_get_state_at(cfg.M, cfg.n, C.c_double(cfg.dt), _p(traj, C.c_float), C.c_double(time), _p(out, C.c_float))
    return out


def shift_traj(cfg: Config, prev) -> np.ndarray:
    prev = _f32(prev); out = np.zeros_like(prev)
    lib().orc_shift_traj(cfg.M, cfg.n, _p(prev, C.c_float), _p(out, C.c_float))
    return out


def const_vel_traj(cfg: Config, pos, vel) -> np.ndarray:
    pos, vel = _f32(pos), _f32(vel); out = np.zeros((cfg.M, cfg.n + 1, 3), np.float32)
    lib().orc_const_vel_traj(cfg.M, cfg.n, C.c_double(cfg.dt), _p(pos, C.c_float), _p(vel, C.c_float), _p(out, C.c_float))
    return out


def is_sol_valid(cfg: Config, ag: Agent, traj, state9, sfc=None) -> bool:
    """TrajPlanner::isSolValid (src/traj_planner.cpp:990-1045)"""
    traj, state9 = _f32(traj), _f32(state9)
    sfc = None if sfc is None else _f32(sfc)
    cc, ca = cfg.c(), ag.c()
    return bool(lib().orc_is_sol_valid(C.byref(cc), C.byref(ca), _p(traj, C.c_float), _p(state9, C.c_float),
                                       _p(sfc, C.c_float) if sfc is not None else None))


# ---------------------------------------------------------------- GoalOptimizer (src/goal_optimizer.cpp)
def goal_rows(cfg: Config, goal, waypoint, lsc_point, lsc_normal, lsc_d, sfc_last=None):
    """rows a t + b >= 0 of the one-variable goal LP (goal_optimizer.cpp:109-165), in the reference's order"""
    goal, waypoint = _f32(goal), _f32(waypoint)
    pt, nr, d = _f32(lsc_point), _f32(lsc_normal), _f64(lsc_d)
    K = pt.shape[0]
    a = np.zeros(2 * cfg.dim + K); b = np.zeros(2 * cfg.dim + K)
    sfc = None if sfc_last is None else _f32(sfc_last)
    cc = cfg.c()
    n = lib().orc_goal_rows(C.byref(cc), _p(goal, C.c_float), _p(waypoint, C.c_float), K, _p(pt, C.c_float),
                            _p(nr, C.c_float), _p(d, C.c_double), _p(sfc, C.c_float) if sfc is not None else None,
                            _p(a, C.c_double), _p(b, C.c_double))
    return a[:n].copy(), b[:n].copy()


def goal_solve(cfg: Config, goal, waypoint, a, b, feas_tol: float = 1e-6):
    """closed-form optimum of the goal LP; returns (goal_out f32[3], t, status) with status 0 ok | 2 infeasible"""
    goal, waypoint = _f32(goal), _f32(waypoint)
    a, b = _f64(a), _f64(b)
    out = np.zeros(3, np.float32); t = C.c_double()
    cc = cfg.c()
    st = lib().orc_goal_solve(C.byref(cc), _p(goal, C.c_float), _p(waypoint, C.c_float), len(a), _p(a, C.c_double),
                              _p(b, C.c_double), C.c_double(feas_tol), _p(out, C.c_float), C.byref(t))
    return out, t.value, st


def goal_solve_highs(a, b):
    """the same LP through HiGHS (scipy.optimize.linprog): the independent solver standing in for CPLEX.
    Returns (t, feasible)."""
    from scipy.optimize import linprog
    a, b = np.asarray(a, float), np.asarray(b, float)
    res = linprog([1.0], A_ub=-a[:, None] if len(a) else None, b_ub=b if len(a) else None,
                  bounds=[(0.0, 1.0 + 1e-5)], method="highs")
    return (float(res.x[0]) if res.status == 0 else float("nan")), res.status == 0


# ---------------------------------------------------------------- HiGHS stand-in for CPLEX
@dataclass
class Solution:
    status: str
    x: np.ndarray
    objective: float          # x'Px + q'x + c0  (== cplex.getObjValue(), traj_optimizer.cpp:100)
    row_dual: np.ndarray      # [ne + ni]
    col_dual: np.ndarray


def solve_highs(qp: QP, tol: float = 1e-9, time_limit: float | None = None) -> Solution:
    """Solve the restated model with HiGHS's QP active-set solver."""
    import scipy
    import scipy.sparse as sp
    from scipy.optimize._highspy import _core as hs
    nv = qp.q.size
    A = sp.csr_matrix(np.vstack([qp.Aeq, qp.G]))
    lo = np.concatenate([qp.beq, qp.rlo]); hi = np.concatenate([qp.beq, qp.rhi])
    lo = np.where(lo <= -INF, -hs.kHighsInf, lo); hi = np.where(hi >= INF, hs.kHighsInf, hi)
    lb = np.where(qp.lb <= -INF, -hs.kHighsInf, qp.lb); ub = np.where(qp.ub >= INF, hs.kHighsInf, qp.ub)
    h = hs._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("primal_feasibility_tolerance", tol)
    h.setOptionValue("dual_feasibility_tolerance", tol)
    if time_limit is not None:
        h.setOptionValue("time_limit", float(time_limit))
    lp = hs.HighsLp()
    lp.num_col_ = nv; lp.num_row_ = A.shape[0]
    lp.col_cost_ = qp.q; lp.col_lower_ = lb; lp.col_upper_ = ub
    lp.row_lower_ = lo; lp.row_upper_ = hi
    lp.offset_ = qp.c0
    lp.a_matrix_.format_ = hs.MatrixFormat.kRowwise
    lp.a_matrix_.start_ = A.indptr; lp.a_matrix_.index_ = A.indices; lp.a_matrix_.value_ = A.data
    h.passModel(lp)
    Hl = sp.csc_matrix(np.tril(2.0 * qp.P))
    hess = hs.HighsHessian()
    hess.dim_ = nv; hess.format_ = hs.HessianFormat.kTriangular
    hess.start_ = Hl.indptr; hess.index_ = Hl.indices; hess.value_ = Hl.data
    h.passHessian(hess)
    h.run()
    sol = h.getSolution()
    x = np.array(sol.col_value)
    status = h.modelStatusToString(h.getModelStatus())
    obj = float(x @ qp.P @ x + qp.q @ x + qp.c0)
    return Solution(status, x, obj, np.array(sol.row_dual), np.array(sol.col_dual))


def solve_dense_ipm(qp: QP, tol: float = 1e-10, max_iter: int = 200) -> Solution:
    """Second, independent checker solver: a textbook dense Mehrotra primal-dual interior-point method on the restated
    model exactly as populatebyrow states it (all variables, equalities kept as equalities, numpy dense KKT solves) --
    no equality elimination, no structure, nothing shared with the CUDA kernel.  Used where HiGHS' active-set QP
    solver reports "Solve error" (it does on some dense communication-range models).  The duals are returned in the
    layout polish() reads (row_dual over [Aeq; G], col_dual over the variable bounds)."""
    nv, ne, ng = qp.q.size, qp.Aeq.shape[0], qp.G.shape[0]
    rows, h, tag = [], [], []
    for i in range(ng):
        if qp.rhi[i] < INF:
            rows.append(qp.G[i]); h.append(qp.rhi[i]); tag.append(("r", i))
        if qp.rlo[i] > -INF:
            rows.append(-qp.G[i]); h.append(-qp.rlo[i]); tag.append(("r", i))
    for j in range(nv):
        if qp.ub[j] < INF:
            e = np.zeros(nv); e[j] = 1; rows.append(e); h.append(qp.ub[j]); tag.append(("c", j))
        if qp.lb[j] > -INF:
            e = np.zeros(nv); e[j] = -1; rows.append(e); h.append(-qp.lb[j]); tag.append(("c", j))
    Gi = np.array(rows); h = np.array(h)
    H = 2.0 * qp.P; A = qp.Aeq; b = qp.beq
    x = np.zeros(nv); y = np.zeros(ne)
    s = np.maximum(h - Gi @ x, 1.0); z = np.ones(len(h))
    status = "Iteration limit"
    for _ in range(max_iter):
        rd = H @ x + qp.q + A.T @ y + Gi.T @ z
        rp = A @ x - b
        rg = Gi @ x + s - h
        mu = float(s @ z) / len(h)
        scale = max(1.0, float(np.abs(qp.q).max()), float(np.abs(H @ x).max()))
        if (np.abs(rd).max() < 1e2 * tol * scale and max(np.abs(rp).max(initial=0.0), np.abs(rg).max()) < 1e2 * tol
                and mu < tol):
            status = "Optimal"
            break
        W = z / s
        K = np.block([[H + Gi.T @ (W[:, None] * Gi), A.T], [A, np.zeros((ne, ne))]])
        lu = np.linalg.inv(K + 1e-14 * np.eye(nv + ne))

        def direction(rc):
            rhs = np.concatenate([-rd + Gi.T @ (rc / s - W * rg), -rp])
            d = lu @ rhs
            dx = d[:nv]
            ds = -rg - Gi @ dx
            dz = -(rc + z * ds) / s
            return dx, d[nv:], ds, dz

        def step(ds, dz):
            a = 1.0
            for v, dv in ((s, ds), (z, dz)):
                neg = dv < 0
                if neg.any():
                    with np.errstate(over="ignore"):
                        a = min(a, float((-v[neg] / dv[neg]).min()))
            return a
        dxa, dya, dsa, dza = direction(s * z)
        aa = step(dsa, dza)
        mu_aff = float((s + aa * dsa) @ (z + aa * dza)) / len(h)
        sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        dx, dy, ds, dz = direction(s * z + dsa * dza - sigma * mu)
        a = min(1.0, 0.995 * step(ds, dz))
        x += a * dx; y += a * dy; s += a * ds; z += a * dz
    row_dual = np.zeros(ne + ng); col_dual = np.zeros(nv)
    row_dual[:ne] = y
    for (kind, i), zi, si in zip(tag, z, s):
        if zi <= si:                                   # inactive at the solution (interior-point active-set indicator)
            continue
        if kind == "r":
            row_dual[ne + i] = max(row_dual[ne + i], zi)
        else:
            col_dual[i] = max(col_dual[i], zi)
    return Solution(status, x, float(x @ qp.P @ x + qp.q @ x + qp.c0), row_dual, col_dual)


def polish(qp: QP, sol: Solution, dual_tol: float = 1e-9, feas_tol: float = 1e-9):
    """Refine a HiGHS solution to ~1e-12: take the rows/bounds HiGHS holds active (non-zero dual),
    keep a linearly independent subset, and solve the equality-constrained KKT system exactly.
    Returns (x, ok); ok is False (and x = sol.x) when the refined point is not primal feasible or
    a multiplier has the wrong sign, i.e. when HiGHS's active set was not the optimal one."""
    import scipy.linalg as sl
    nv = qp.q.size
    ne = qp.Aeq.shape[0]
    rows, rhs, sign = [], [], []
    rd = sol.row_dual[ne:]
    ax = qp.G @ sol.x
    for i in np.where(np.abs(rd) > dual_tol)[0]:
        # pick the side the point sits on
        if qp.rhi[i] < INF and (qp.rlo[i] <= -INF or abs(ax[i] - qp.rhi[i]) <= abs(ax[i] - qp.rlo[i])):
            rows.append(qp.G[i]); rhs.append(qp.rhi[i]); sign.append(+1)
        else:
            rows.append(qp.G[i]); rhs.append(qp.rlo[i]); sign.append(-1)
    for j in np.where(np.abs(sol.col_dual) > dual_tol)[0]:
        e = np.zeros(nv); e[j] = 1
        if abs(sol.x[j] - qp.ub[j]) <= abs(sol.x[j] - qp.lb[j]):
            rows.append(e); rhs.append(qp.ub[j]); sign.append(+1)
        else:
            rows.append(e); rhs.append(qp.lb[j]); sign.append(-1)
    Aall = np.vstack([qp.Aeq] + ([np.array(rows)] if rows else []))
    ball = np.concatenate([qp.beq, np.array(rhs)])
    _, R, piv = sl.qr(Aall.T, pivoting=True, mode="economic")
    dg = np.abs(np.diag(R))
    rk = int((dg > 1e-10 * dg[0]).sum())
    keep = np.sort(piv[:rk])
    Ak, bk = Aall[keep], ball[keep]
    KKT = np.block([[2.0 * qp.P, Ak.T], [Ak, np.zeros((rk, rk))]])
    try:
        z = np.linalg.solve(KKT, np.concatenate([-qp.q, bk]))
    except np.linalg.LinAlgError:
        return sol.x, False
    x = z[:nv]
    gx = qp.G @ x
    feas = (np.all(gx <= qp.rhi + feas_tol) and np.all(gx >= qp.rlo - feas_tol)
            and np.all(x <= qp.ub + feas_tol) and np.all(x >= qp.lb - feas_tol))
    mult = z[nv:]
    ok_sign = True
    for pos, r in enumerate(keep):
        if r >= ne and mult[pos] * sign[r - ne] < -1e-7 * max(1.0, np.abs(mult).max()):
            ok_sign = False
    if feas and ok_sign and np.abs(x - sol.x).max() < 1e-3:
        return x, True
    return sol.x, False


# ---------------------------------------------------------------- KKT certificate
def kkt_certificate(qp: QP, x: np.ndarray) -> dict:
    """Solver-independent optimality certificate for a primal point x of the restated model.

    Every inequality (ranged rows and finite variable bounds) is written a.x <= h.  Multipliers
    are recovered by non-negative least squares on the rows that are active within `act_tol`,
    so the certificate needs no dual output from the solver under test:
        stationarity  || 2Px + q + Aeq' y + Gact' z ||_inf   (z >= 0)
        primal        max(|Aeq x - beq|, (a.x - h)+)
    """
    from scipy.optimize import nnls
    nv = x.size
    rows, rhs = [], []
    for i in range(qp.G.shape[0]):
        if qp.rhi[i] < INF:
            rows.append(qp.G[i]); rhs.append(qp.rhi[i])
        if qp.rlo[i] > -INF:
            rows.append(-qp.G[i]); rhs.append(-qp.rlo[i])
    for j in range(nv):
        if qp.ub[j] < INF:
            e = np.zeros(nv); e[j] = 1; rows.append(e); rhs.append(qp.ub[j])
        if qp.lb[j] > -INF:
            e = np.zeros(nv); e[j] = -1; rows.append(e); rhs.append(-qp.lb[j])
    Gi = np.array(rows); hi = np.array(rhs)
    norms = np.maximum(1.0, np.linalg.norm(Gi, axis=1))
    slack = (hi - Gi @ x) / norms
    prim_ineq = float(max(0.0, -slack.min())) if slack.size else 0.0
    eqn = np.maximum(1.0, np.linalg.norm(qp.Aeq, axis=1))
    prim_eq = float(np.abs((qp.Aeq @ x - qp.beq) / eqn).max())
    grad = 2.0 * qp.P @ x + qp.q
    act = np.where(slack <= 1e-6)[0]
    # eliminate the free equality multipliers by projecting on null(Aeq)
    U, S, Vt = np.linalg.svd(qp.Aeq, full_matrices=True)
    rank = int((S > 1e-10 * S[0]).sum())
    Z = Vt[rank:].T
    gz = Z.T @ grad
    if act.size:
        Gz = (Gi[act] / norms[act, None]) @ Z
        z, _ = nnls(Gz.T, -gz, maxiter=50 * Gz.shape[0] + 500)
        res = gz + Gz.T @ z
    else:
        z = np.zeros(0); res = gz
    return {"stationarity": float(np.abs(res).max()), "primal_eq": prim_eq, "primal_ineq": prim_ineq,
            "n_active": int(act.size), "max_multiplier": float(z.max()) if z.size else 0.0,
            "grad_scale": float(np.abs(grad).max())}


# ------------------------------------------------------------------------------------------------
# Safe Flight Corridor construction (oracle/sfc_oracle.c).  PARITY UNPINNED: octomap / dynamicEDT3D absent.
class Map:
    """occupancy grid + nearest-obstacle queries of a world CSV (MapManager::updateOctreeFromCSV + DynamicEDTOctomap)"""

    def __init__(self, boxes, world_min, world_max, resolution: float = 0.1, max_dist: float = 1.0):
        self.boxes = np.ascontiguousarray(np.asarray(boxes, np.float64).reshape(-1, 6))
        self.world_min = np.asarray(world_min, np.float32); self.world_max = np.asarray(world_max, np.float32)
        self.res = float(resolution)
        self.h = C.c_void_p(lib().orc_map_build(_p(self.boxes, C.c_double), C.c_int(self.boxes.shape[0]), C.c_double(self.res),
                                                _p(self.world_min, C.c_float), _p(self.world_max, C.c_float), C.c_double(max_dist)))
        n3 = (C.c_int * 3)(); k0 = (C.c_int * 3)()
        self.maxd2 = lib().orc_map_dims(self.h, n3, k0)
        self.n = tuple(n3); self.key0 = tuple(k0)

    def __del__(self):
        try:
            lib().orc_map_free(self.h)
        except Exception:
            pass

    def occupancy(self) -> np.ndarray:
        ptr = lib().orc_map_occupancy(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=self.n).copy()

    def closest(self) -> np.ndarray:
        """[nx, ny, nz, 3] nearest occupied cell of every cell (-1: none within max_dist); slow (brute force)"""
        ptr = lib().orc_map_closest(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int)), shape=self.n + (3,)).copy()

    def is_obstacle_in_sfc(self, box, margin: float) -> bool:
        b = np.ascontiguousarray(box, np.float32).reshape(6)
        return bool(lib().orc_is_obstacle_in_sfc(self.h, _p(b, C.c_float), C.c_double(margin)))

    def expand_sfc(self, initial, margin: float, goal=None):
        """(ok, box): CollisionConstraints::expandSFC, goal-directed axis order when `goal` is given"""
        b = np.ascontiguousarray(initial, np.float32).reshape(6); out = np.zeros(6, np.float32)
        g = None if goal is None else np.ascontiguousarray(goal, np.float32)
        ok = lib().orc_expand_sfc(self.h, _p(b, C.c_float), None if g is None else _p(g, C.c_float), C.c_double(margin), _p(out, C.c_float))
        return bool(ok), out

    def sfc_initialize(self, position, radius: float):
        pos = np.ascontiguousarray(position, np.float32); out = np.zeros(6, np.float32)
        ok = lib().orc_sfc_initialize(self.h, _p(pos, C.c_float), C.c_double(radius), _p(out, C.c_float))
        return bool(ok), out

    def sfc_from_point(self, point, goal, prev, radius: float):
        pt = np.ascontiguousarray(point, np.float32); g = np.ascontiguousarray(goal, np.float32)
        pv = np.ascontiguousarray(prev, np.float32).reshape(6); out = np.zeros(6, np.float32)
        ok = lib().orc_sfc_from_point(self.h, _p(pt, C.c_float), _p(g, C.c_float), _p(pv, C.c_float), C.c_double(radius), _p(out, C.c_float))
        return int(ok), out

    def sfc_from_convex_hull(self, hull2, next_waypoint, prev, radius: float):
        h2 = np.ascontiguousarray(hull2, np.float32).reshape(2, 3); w = np.ascontiguousarray(next_waypoint, np.float32)
        pv = np.ascontiguousarray(prev, np.float32).reshape(6); out = np.zeros(6, np.float32)
        ok = lib().orc_sfc_from_convex_hull(self.h, _p(h2, C.c_float), _p(w, C.c_float), _p(pv, C.c_float), C.c_double(radius), _p(out, C.c_float))
        return int(ok), out


# ------------------------------------------------------------------------------------------------
# Interior-point CPU solver (oracle/pdip_cpu.c): third checker solver and the same-algorithm-class CPU baseline
def solve_pdip_c(qp: "QP", tol: float = 1e-10, max_iter: int = 100) -> "Solution":
    """Mehrotra PDIP in C on the restated model as it stands (equalities kept, quasi-definite KKT + LDL')"""
    nv, ne, ng = qp.q.size, qp.Aeq.shape[0], qp.G.shape[0]
    x = np.zeros(nv); info = np.zeros(4)
    arr = lambda a: np.ascontiguousarray(a, np.float64)
    P, q, A, b, G, rlo, rhi, lb, ub = (arr(v) for v in (qp.P, qp.q, qp.Aeq, qp.beq, qp.G, qp.rlo, qp.rhi, qp.lb, qp.ub))
    st = pdip_lib().orc_pdip_solve(nv, ne, ng, _p(P, C.c_double), _p(q, C.c_double), _p(A, C.c_double), _p(b, C.c_double),
                                   _p(G, C.c_double), _p(rlo, C.c_double), _p(rhi, C.c_double), _p(lb, C.c_double),
                                   _p(ub, C.c_double), C.c_double(tol), max_iter, _p(x, C.c_double), _p(info, C.c_double))
    status = {0: "Optimal", 1: "Iteration limit", 3: "Numerical"}[st]
    sol = Solution(status, x, float(x @ qp.P @ x + qp.q @ x + qp.c0), np.zeros(ne + ng), np.zeros(nv))
    sol.iterations = int(info[0])
    return sol


def replan_batch_pdip(cfg: "Config", generator: int, agents: list, own_traj, obs_offsets, obs_index, radius, downwash, goal,
                      position, threads: int = 0):
    """The reference's per-agent path (LSC generation + model build + interior-point solve) for a whole batch in C with
    OpenMP over the agents.  Returns (ctrl [n, nv], status [n], iters [n], seconds of the parallel region)."""
    n = len(agents)
    nv = cfg.dim * cfg.M * (cfg.n + 1)
    cc = cfg.c()
    ags = (_Agent * n)(*[a.c() for a in agents])
    dw = np.ascontiguousarray(downwash, np.float64)
    own = np.ascontiguousarray(own_traj, np.float32); off = np.ascontiguousarray(obs_offsets, np.int32); idx = np.ascontiguousarray(obs_index, np.int32)
    rf = np.ascontiguousarray(radius, np.float32); df = np.ascontiguousarray(downwash, np.float32)
    g = np.ascontiguousarray(goal, np.float32); pos = np.ascontiguousarray(position, np.float32)
    ctrl = np.zeros((n, nv)); status = np.zeros(n, np.int32); iters = np.zeros(n, np.int32)
    sec = pdip_lib().orc_replan_batch_pdip(C.byref(cc), generator, n, ags, _p(dw, C.c_double), _p(own, C.c_float), _p(off, C.c_int),
                                           _p(idx, C.c_int), _p(rf, C.c_float), _p(df, C.c_float), _p(g, C.c_float),
                                           _p(pos, C.c_float), _p(ctrl, C.c_double), _p(status, C.c_int), _p(iters, C.c_int), threads)
    return ctrl, status, iters, float(sec)


# ------------------------------------------------------------------------------------------------
# CPLEX-LP files of the restated models: the format cplex.exportModel("QPmodel_trajOpt.lp") writes when param.log_solver is
# set (src/traj_optimizer.cpp:45-49).  Anyone with CPLEX can solve tests/golden/lp/*.lp and close the pin on the fixture
# solutions; HiGHS reads the same files (read_lp_highs), which is how the writer itself is tested.
def variable_names(M: int, n: int, dim: int) -> list[str]:
    """x_m_i / y_m_i / z_m_i in the order of populatebyrow (traj_optimizer.cpp:238-270)"""
    return [f"{'xyz'[k]}_{m}_{i}" for k in range(dim) for m in range(M) for i in range(n + 1)]


def _lp_num(v: float) -> str:
    return repr(float(v))


def write_lp(qp: QP, path: str, names: list[str] | None = None, comment: str = "") -> None:
    """min x'Px + q'x + c0  s.t.  Aeq x = beq, rlo <= G x <= rhi, lb <= x <= ub as a CPLEX LP file.  The quadratic part is
    written in the format's [ ... ] / 2 bracket (twice the entries of P, both triangles merged)."""
    nv = qp.q.size
    names = names or [f"v{j}" for j in range(nv)]
    out = []
    if comment:
        out += ["\\ " + line for line in comment.splitlines()]
    out.append("Minimize")
    lin = " ".join(f"{'+' if qp.q[j] >= 0 else '-'} {_lp_num(abs(qp.q[j]))} {names[j]}" for j in range(nv) if qp.q[j] != 0.0)
    out.append(" obj: " + (lin if lin else f"0 {names[0]}"))
    quad = []
    for i in range(nv):
        if qp.P[i, i] != 0.0:
            quad.append(f"{'+' if qp.P[i, i] >= 0 else '-'} {_lp_num(abs(2.0 * qp.P[i, i]))} {names[i]} ^2")
        for j in range(i + 1, nv):
            c = 2.0 * (qp.P[i, j] + qp.P[j, i])
            if c != 0.0:
                quad.append(f"{'+' if c >= 0 else '-'} {_lp_num(abs(c))} {names[i]} * {names[j]}")
    for k in range(0, len(quad), 6):
        out.append("   " + ("+ [ " if k == 0 else "") + " ".join(quad[k:k + 6]))
    if quad:
        out.append("   ] / 2")
    if qp.c0 != 0.0:                                      # constant of the terminal cost (part of getObjValue, :100)
        out.append(f"   {'+' if qp.c0 >= 0 else '-'} {_lp_num(abs(qp.c0))} objconst")
    out.append("Subject To")

    def row(coefs):
        nz = np.nonzero(coefs)[0]
        return " ".join(f"{'+' if coefs[j] >= 0 else '-'} {_lp_num(abs(coefs[j]))} {names[j]}" for j in nz) if len(nz) else f"0 {names[0]}"
    for e in range(qp.Aeq.shape[0]):
        out.append(f" e{e}: {row(qp.Aeq[e])} = {_lp_num(qp.beq[e])}")
    for r in range(qp.G.shape[0]):
        lo, hi = qp.rlo[r], qp.rhi[r]
        if lo > -INF and hi < INF:
            out.append(f" c{r}: {row(qp.G[r])} >= {_lp_num(lo)}")
            out.append(f" c{r}u: {row(qp.G[r])} <= {_lp_num(hi)}")
        elif hi < INF:
            out.append(f" c{r}: {row(qp.G[r])} <= {_lp_num(hi)}")
        else:
            out.append(f" c{r}: {row(qp.G[r])} >= {_lp_num(lo)}")
    out.append("Bounds")
    for j in range(nv):
        lo, hi = qp.lb[j], qp.ub[j]
        if lo <= -INF and hi >= INF:
            out.append(f" {names[j]} free")
        else:
            out.append(f" {'-inf' if lo <= -INF else _lp_num(lo)} <= {names[j]} <= {'+inf' if hi >= INF else _lp_num(hi)}")
    if qp.c0 != 0.0:
        out.append(" objconst = 1")
    out.append("End")
    opener = __import__("gzip").open if path.endswith(".gz") else open
    with opener(path, "wt") as f:
        f.write("\n".join(out) + "\n")


def solve_lp_file_highs(path: str, tol: float = 1e-9, time_limit: float | None = 60.0):
    """read a CPLEX-LP file with HiGHS (Highs::readModel) and solve it; returns (status, {name: value}, objective)"""
    import gzip, shutil, tempfile
    from scipy.optimize._highspy import _core as hs
    tmp = None
    if path.endswith(".gz"):
        tmp = tempfile.NamedTemporaryFile(suffix=".lp", delete=False)
        with gzip.open(path, "rb") as f:
            shutil.copyfileobj(f, tmp)
        tmp.close(); path = tmp.name
    h = hs._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("primal_feasibility_tolerance", tol); h.setOptionValue("dual_feasibility_tolerance", tol)
    if time_limit:
        h.setOptionValue("time_limit", float(time_limit))
    st = h.readModel(path)
    h.run()
    sol = h.getSolution()
    lp = h.getLp()
    names = list(lp.col_names_)
    vals = dict(zip(names, sol.col_value))
    status = h.modelStatusToString(h.getModelStatus())
    obj = h.getInfo().objective_function_value
    if tmp:
        os.unlink(tmp.name)
    return status, vals, obj
