"""CPU tests of the oracle: known-answer constants, golden vectors from the reference's own openGJK,
model sizes, and the HiGHS-based solution of the restated QP."""
import os

import numpy as np
import pytest

from common import near_goals, oracle_agent, oracle_config, oracle_lsc, oracle_planes, oracle_qp_from_planes, oracle_solution
from lsc_dr_planner_b200 import workloads as W
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_bernstein_basis_known_answer():
    """SURVEY.md appendix B / include/polynomial.hpp:281-294"""
    B = orc.bernstein_basis(5)
    want = np.array([[1, -5, 10, -10, 5, -1], [0, 5, -20, 30, -20, 5], [0, 0, 10, -30, 30, -10],
                     [0, 0, 0, 10, -20, 10], [0, 0, 0, 0, 5, -5], [0, 0, 0, 0, 0, 1]], float)
    assert np.array_equal(B, want)


def test_qbase_known_answer():
    """Q_base * dt^5 is the integer matrix of SURVEY.md appendix B (src/traj_optimizer.cpp:163-178)"""
    want = np.array([[720, -1800, 1200, 0, 0, -120], [-1800, 4800, -3600, 0, 600, 0], [1200, -3600, 3600, -1200, 0, 0],
                     [0, 0, -1200, 3600, -3600, 1200], [0, 600, 0, -3600, 4800, -1800], [-120, 0, 0, 1200, -1800, 720]], float)
    for dt in (0.2, 0.5):
        Q = orc.qbase(5, 3, 1, dt)
        assert np.allclose(Q * dt ** 5, want, rtol=1e-12, atol=1e-9)
        assert np.allclose(Q.sum(axis=1), 0, atol=1e-6 * np.abs(Q).max())
        assert np.linalg.matrix_rank(Q, tol=1e-8 * np.abs(Q).max()) == 3
    ev = np.sort(np.linalg.eigvalsh(orc.qbase(5, 3, 1, 0.2)))
    assert np.allclose(ev[3:], [3.63e6, 2.625e7, 2.712e7], rtol=2e-3)


def test_aeq_base_requires_n5_phi3():
    with pytest.raises(ValueError):
        orc.aeq_base(5, n=4)
    A = orc.aeq_base(5)
    assert A.shape == (9, 30)
    # C0 row between segments 1 and 2: last point of segment 1 minus first of segment 2
    assert A[0, 11] == 1 and A[0, 12] == -1


@pytest.mark.parametrize("M,dim,mode,K,use_sfc,comm", [(5, 3, 1, 40, False, 0.0), (10, 2, 1, 9, True, 3.0), (5, 3, 0, 7, False, 0.0)])
def test_model_sizes_follow_survey_formulas(M, dim, mode, K, use_sfc, comm):
    """eq = D(3M+2) [LSC] or 3DM; SFC 2D(6M-3); LSC K(6M-3); vel 2D(5M-2); acc 2D(4M-1)"""
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=mode, use_sfc=use_sfc, comm_range=comm)
    batch = W.make_forest_batch(64, K=K, cfg=cfg)
    pt, nr, d = oracle_lsc(batch, 0, orc.GEN_LSC)
    cfgo = oracle_config(batch.cfg)
    sfc = np.tile(np.array([-50, -50, 0, 50, 50, 3], np.float32), (M, 1)) if use_sfc else None
    qp = orc.qp_build(cfgo, oracle_agent(batch, 0), pt, nr, d, sfc)
    assert qp.q.size == dim * 6 * M
    assert qp.Aeq.shape[0] == (dim * (3 * M + 2) if mode == 1 else 3 * dim * M)
    assert np.linalg.matrix_rank(qp.Aeq) == qp.Aeq.shape[0]
    rows = K * (6 * M - 3) + 2 * dim * (5 * M - 2) + 2 * dim * (4 * M - 1)
    if use_sfc:
        rows += 2 * dim * (6 * M - 3)
    if comm > 0:
        rows += dim * M * (M + 1) + 2 * dim * M
    assert qp.G.shape[0] == rows
    # the first three control points of segment 0 are free variables (traj_optimizer.cpp:260-263)
    for k in range(dim):
        assert np.all(qp.lb[k * 6 * M:k * 6 * M + 3] <= -1e29) and np.all(qp.ub[k * 6 * M + 3:(k + 1) * 6 * M] < 1e29)


def test_terminal_segments_rule():
    """getTerminalSegments_old, src/traj_optimizer.cpp:530-538"""
    cfg = orc.Config()
    for dist, want in ((10.0, 1), (0.79, 1), (0.61, 1), (0.59, 2), (0.05, 4), (0.0, 5)):
        ag = orc.Agent(np.zeros(3), np.zeros(3), np.zeros(3), np.array([dist, 0, 0]))
        assert orc.terminal_segments(cfg, ag) == want, dist


def test_min_norm_hull_against_reference_gjk_golden():
    """orc_min_norm_hull reproduces the reference's openGJK witness vector on the committed golden cases"""
    g = np.load(os.path.join(GOLDEN, "gjk_golden.npz"))
    worst = 0.0
    for pts, v_ref, d_ref in zip(g["pts"], g["v"], g["dist"]):
        v, d = orc.min_norm_hull(pts)
        worst = max(worst, np.abs(v - v_ref).max())
        assert abs(d - d_ref) <= 1e-7 * max(1.0, d_ref)
    assert worst < 1e-7          # openGJK's own exit tolerance is 1e-10 relative on |v|^2


@pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref not built (reference sources absent)")
def test_min_norm_hull_against_live_reference_gjk():
    rng = np.random.default_rng(7)
    for _ in range(2000):
        pts = (rng.normal(size=(6, 3)) * rng.uniform(0.01, 2) + rng.normal(size=3) * rng.uniform(0, 3)).astype(np.float32).astype(np.float64)
        v, d = orc.min_norm_hull(pts)
        v_ref, d_ref = orc.ref_gjk(pts)
        assert np.abs(v - v_ref).max() < 1e-7


def test_lsc_rule_initial_trajectory_is_feasible():
    """the own initial trajectory satisfies every LSC row it generates with margin
    (dist - R)/2 >= 0, for LSC, CLSC and BVC"""
    for gen, M, dim in ((orc.GEN_LSC, 5, 3), (orc.GEN_CLSC, 10, 2), (orc.GEN_BVC, 5, 3)):
        cfg = W.PlannerConfig(M=M, dim=dim)
        batch = W.make_forest_batch(64, K=12, cfg=cfg)
        if gen == orc.GEN_CLSC:
            near_goals(batch)
        for a in (0, 7, 33):
            pt, nr, d = oracle_lsc(batch, a, gen)
            c = batch.own_traj[a].astype(np.float64)
            margin = np.einsum("kmid,kmid->kmi", nr.astype(np.float64)[..., :dim], (c[None] - pt.astype(np.float64))[..., :dim]) - d
            if gen != orc.GEN_BVC:
                assert margin.min() > 0, (gen, a, margin.min())
            assert np.all(np.abs(np.linalg.norm(nr[:, :, 0, :2] if dim == 2 else nr[:, :, 0, :] * [1, 1, 2.0], axis=-1) - 1) < 1e-5)


def test_segment_closest_points_cases():
    """closestPointsBetweenLineSegments (include/geometry.hpp:174-264): crossing, parallel, degenerate"""
    cp1, cp2, d = orc.closest_points_segments([0, 0, 0], [1, 0, 0], [0.5, -1, 1], [0.5, 1, 1])
    assert abs(d - 1) < 1e-6 and np.allclose(cp1, [0.5, 0, 0], atol=1e-6) and np.allclose(cp2, [0.5, 0, 1], atol=1e-6)
    cp1, cp2, d = orc.closest_points_segments([0, 0, 0], [1, 0, 0], [2, 1, 0], [3, 1, 0])
    assert abs(d - np.sqrt(2)) < 1e-6
    cp1, cp2, d = orc.closest_points_segments([0, 0, 0], [0, 0, 0], [1, -1, 0], [1, 1, 0])
    assert abs(d - 1) < 1e-6 and np.allclose(cp2, [1, 0, 0], atol=1e-6)


def test_closed_loop_glue():
    """getStateAt at t=0 returns the initial state; the shift collapses the last segment (traj_planner.cpp:402-411)"""
    cfg = W.PlannerConfig()
    batch = W.make_forest_batch(64, K=4, cfg=cfg)
    cfgo = oracle_config(cfg)
    for a in (0, 9):
        st = orc.get_state_at(cfgo, batch.own_traj[a], 0.0)
        assert np.allclose(st, batch.state[a], rtol=1e-4, atol=2e-4)
        sh = orc.shift_traj(cfgo, batch.own_traj[a])
        assert np.array_equal(sh[:-1], batch.own_traj[a][1:])
        assert np.array_equal(sh[-1], np.tile(batch.own_traj[a][-1, -1], (6, 1)))
        end = orc.get_state_at(cfgo, batch.own_traj[a], cfg.M * cfg.dt)
        assert np.allclose(end[:3], batch.own_traj[a][-1, -1], atol=1e-6) and np.allclose(end[3:], 0, atol=1e-4)
    cv = orc.const_vel_traj(cfgo, [1, 2, 3], [0.5, 0, -0.5])
    assert np.allclose(cv[2, 3], np.array([1, 2, 3]) + np.array([0.5, 0, -0.5]) * (2 * 6 + 3) * 0.04, atol=1e-6)


def test_qp_golden_solutions_are_kkt_points():
    """the committed polished solutions satisfy the KKT conditions of the freshly restated model"""
    g = np.load(os.path.join(GOLDEN, "qp_golden.npz"))
    cases = {"m5d3lsc": (W.PlannerConfig(M=5, dim=3, planner_mode=1), 40), "m10d2lsc": (W.PlannerConfig(M=10, dim=2, planner_mode=1), 9),
             "m5d3dlsc": (W.PlannerConfig(M=5, dim=3, planner_mode=0), 16)}
    for name, (cfg, K) in cases.items():
        batch = W.make_forest_batch(64, K=K, cfg=cfg)
        off, normals, rhs = g[name + "_off"], g[name + "_normals"], g[name + "_rhs"]
        off2, n2, r2 = oracle_planes(batch, [0, 5, 11, 17, 23, 42], orc.GEN_LSC)
        assert np.array_equal(off, off2) and np.allclose(normals, n2, atol=0) and np.allclose(rhs, r2, atol=1e-12)
        for x, a, i in zip(g[name + "_x"][:3], g[name + "_agents"][:3], g[name + "_keep"][:3]):
            qp = oracle_qp_from_planes(batch, int(a), normals[off[i]:off[i + 1]], rhs[off[i]:off[i + 1]])
            cert = orc.kkt_certificate(qp, x)
            assert cert["primal_eq"] < 1e-9 and cert["primal_ineq"] < 1e-9
            assert cert["stationarity"] < 1e-6 * max(1.0, cert["grad_scale"]), cert


def test_highs_solution_matches_golden():
    g = np.load(os.path.join(GOLDEN, "qp_golden.npz"))
    cfg = W.PlannerConfig(M=5, dim=3, planner_mode=1)
    batch = W.make_forest_batch(64, K=40, cfg=cfg)
    off, normals, rhs = g["m5d3lsc_off"], g["m5d3lsc_normals"], g["m5d3lsc_rhs"]
    a, i = int(g["m5d3lsc_agents"][0]), int(g["m5d3lsc_keep"][0])
    qp = oracle_qp_from_planes(batch, a, normals[off[i]:off[i + 1]], rhs[off[i]:off[i + 1]])
    x, ok = oracle_solution(qp)
    assert ok and np.abs(x - g["m5d3lsc_x"][0]).max() < 1e-8


def test_interior_point_c_solver_matches_the_golden_solutions():
    """oracle/pdip_cpu.c (third, independent solver; also bench.py's same-algorithm-class CPU baseline) on the committed
    golden QPs, and its batched OpenMP driver on a small batch"""
    g = np.load(os.path.join(GOLDEN, "qp_golden.npz"))
    for name, cfg, K in (("m5d3lsc", W.PlannerConfig(M=5, dim=3, planner_mode=1), 40), ("m10d2lsc", W.PlannerConfig(M=10, dim=2, planner_mode=1), 9)):
        batch = W.make_forest_batch(64, K=K, cfg=cfg)
        off, normals, rhs = g[name + "_off"], g[name + "_normals"], g[name + "_rhs"]
        agents = [0, 5, 11, 17, 23, 42]
        for x, i in zip(g[name + "_x"][:2], [int(v) for v in g[name + "_keep"][:2]]):
            qp = oracle_qp_from_planes(batch, agents[i], normals[off[i]:off[i + 1]], rhs[off[i]:off[i + 1]])
            sol = orc.solve_pdip_c(qp)
            assert sol.status == "Optimal" and np.abs(sol.x - x).max() < 1e-6, (name, i, sol.status, np.abs(sol.x - x).max())
    cfg = W.PlannerConfig()
    batch = W.make_forest_batch(32, K=12, cfg=cfg)
    ctrl, status, iters, sec = orc.replan_batch_pdip(oracle_config(cfg), orc.GEN_LSC, [oracle_agent(batch, a) for a in range(32)],
                                                     batch.own_traj, batch.obs_offsets, batch.obs_index, batch.agent_meta[:, 0],
                                                     batch.agent_meta[:, 1], batch.goal, batch.state[:, :3], threads=2)
    assert (status == 0).mean() > 0.9 and sec > 0
    for a in (0, 7):
        o, n_, r_ = oracle_planes(batch, [a], orc.GEN_LSC)
        xe, ok = oracle_solution(oracle_qp_from_planes(batch, a, n_, r_))
        if ok and status[a] == 0:
            assert np.abs(ctrl[a] - xe).max() < 1e-6
