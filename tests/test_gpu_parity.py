"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on identical inputs."""
import numpy as np
import pytest

from common import (near_goals, oracle_config, oracle_lsc, oracle_planes, oracle_qp_from_planes, oracle_solution)
from lsc_dr_planner_b200 import capi
from lsc_dr_planner_b200 import workloads as W
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

CTRL_TOL = 1e-5        # metres, SURVEY.md 8(c): control points vs the oracle's polished optimum
OBJ_RTOL = 1e-6        # relative objective difference
PRIMAL_TOL = 1e-8      # scaled primal infeasibility of the returned point


def _planner(cfg):
    from lsc_dr_planner_b200.planner import BatchPlanner
    return BatchPlanner(cfg, device=0)


def _solve_host(planner, batch, agents, off, normals, rhs, sfc=None, want_dual=False, warm=False):
    n = len(agents)
    cfg = batch.cfg
    state = np.ascontiguousarray(batch.state[agents]); goal = np.ascontiguousarray(batch.goal[agents])
    limits = np.ascontiguousarray(batch.limits[agents])
    ctrl = np.zeros((n, cfg.dim * cfg.M * 6)); cost = np.zeros(n); status = np.zeros(n, np.int32)
    iters = np.zeros(n, np.int32); kkt = np.zeros((n, 4))
    dual = np.zeros((n, planner.qp.dual_stride)) if want_dual else None
    initial_traj = np.ascontiguousarray(batch.own_traj[agents]) if warm else None
    planner.qp.solve_host(n, state, goal, limits, sfc, off, normals, rhs, ctrl, cost, status, iters, kkt, dual,
                          initial_traj=initial_traj)
    return ctrl, cost, status, iters, kkt, dual


def _check_against_oracle(batch, agents, off, normals, rhs, ctrl, cost, status, sfc=None, min_checked=1):
    checked = 0
    errs = []
    for i, a in enumerate(agents):
        sl = slice(off[i], off[i + 1])
        qp = oracle_qp_from_planes(batch, a, normals[sl], rhs[sl], None if sfc is None else sfc[i])
        cert = orc.kkt_certificate(qp, ctrl[i])
        assert cert["primal_eq"] < PRIMAL_TOL, (a, cert)
        assert cert["primal_ineq"] < PRIMAL_TOL, (a, cert)
        xe, ok = oracle_solution(qp)
        if not ok:
            continue
        checked += 1
        assert status[i] == capi.STATUS_OK, (a, status[i])
        err = np.abs(ctrl[i] - xe).max()
        errs.append(err)
        assert err < CTRL_TOL, (a, err)
        obj = xe @ qp.P @ xe + qp.q @ xe + qp.c0
        assert abs(cost[i] - obj) <= OBJ_RTOL * max(1.0, abs(obj)), (a, cost[i], obj)
    assert checked >= min_checked
    return np.array(errs)


PDIP_ONLY = 9          # lscqp_config.presolve: presolve on (bit 0), dual active-set first pass off (bit 3)


@pytest.mark.parametrize("solver", ["das", "pdip"])
@pytest.mark.parametrize("M,dim,mode,K", [(5, 3, capi.MODE_LSC, 40), (5, 3, capi.MODE_DLSC, 40), (10, 2, capi.MODE_LSC, 9),
                                          (5, 2, capi.MODE_LSC, 12), (10, 3, capi.MODE_DLSC, 40), (5, 3, capi.MODE_BVC, 7)])
def test_solve_parity_real_rule(M, dim, mode, K, solver):
    """config 2 shape (and the launch-file shape M=10/D=2): oracle-generated planes, GPU solve vs oracle optimum; once
    through the default dispatch (dual active-set first pass, interior point for what it defers) and once through the
    interior-point instances alone"""
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=mode, presolve=1 if solver == "das" else PDIP_ONLY)
    batch = W.make_forest_batch(64, K=K, cfg=cfg)
    agents = list(range(0, 64, 4))
    gen = orc.GEN_BVC if mode == capi.MODE_BVC else orc.GEN_LSC
    off, normals, rhs = oracle_planes(batch, agents, gen)
    planner = _planner(batch.cfg)
    ctrl, cost, status, iters, kkt, _ = _solve_host(planner, batch, agents, off, normals, rhs)
    errs = _check_against_oracle(batch, agents, off, normals, rhs, ctrl, cost, status, min_checked=8)
    assert np.median(errs) < 1e-7
    # interior point: a few dozen iterations at most; active set: one row enters or leaves per iteration, capped at 4 NR + 40
    assert iters.max() < (40 if solver == "pdip" else 4 * dim * 3 * M + 40)


@pytest.mark.parametrize("M,dim,mode,K", [(5, 3, capi.MODE_LSC, 40), (10, 2, capi.MODE_LSC, 9), (5, 3, capi.MODE_DLSC, 24)])
def test_warm_start_parity(M, dim, mode, K):
    """initial_traj as the interior-point solver's starting point: same optimum, fewer iterations than the cold start
    (the dual active-set pass starts from the unconstrained minimiser and ignores it: switched off here)"""
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=mode, presolve=PDIP_ONLY)
    batch = W.make_forest_batch(64, K=K, cfg=cfg)
    agents = list(range(1, 64, 4))
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    planner = _planner(batch.cfg)
    ctrl, cost, status, iters, kkt, _ = _solve_host(planner, batch, agents, off, normals, rhs, warm=True)
    _check_against_oracle(batch, agents, off, normals, rhs, ctrl, cost, status, min_checked=8)
    ctrl_c, cost_c, status_c, iters_c, _, _ = _solve_host(planner, batch, agents, off, normals, rhs, warm=False)
    assert np.abs(ctrl - ctrl_c).max() < CTRL_TOL
    assert iters.mean() < iters_c.mean()


def test_warm_start_from_infeasible_trajectory():
    """a starting trajectory that violates rows (and the model's equalities) is only a hint: same optimum"""
    batch = W.make_forest_batch(64, K=40)
    agents = list(range(0, 64, 8))
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    planner = _planner(batch.cfg)
    rng = np.random.default_rng(3)
    batch.own_traj = (batch.own_traj + rng.normal(0, 0.5, batch.own_traj.shape)).astype(np.float32)
    ctrl, cost, status, iters, kkt, _ = _solve_host(planner, batch, agents, off, normals, rhs, warm=True)
    _check_against_oracle(batch, agents, off, normals, rhs, ctrl, cost, status, min_checked=6)


def test_presolve_is_exact():
    """dropping obstacles proven inactive by bound propagation does not move the optimum"""
    import copy
    batch = W.make_forest_batch(256, K=40)
    agents = list(range(0, 256, 8))
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    cfg_off = copy.copy(batch.cfg); cfg_off.presolve = False
    p_on, p_off = _planner(batch.cfg), _planner(cfg_off)
    c_on, cost_on, st_on, it_on, _, dual_on = _solve_host(p_on, batch, agents, off, normals, rhs, want_dual=True)
    c_off, cost_off, st_off, it_off, _, dual_off = _solve_host(p_off, batch, agents, off, normals, rhs, want_dual=True)
    assert (st_on == 0).all() and (st_off == 0).all()
    assert np.abs(c_on - c_off).max() < CTRL_TOL
    assert np.abs(cost_on - cost_off).max() <= OBJ_RTOL * np.abs(cost_off).max()
    # multipliers of dropped rows are exactly zero with presolve and negligible without (the others need not be
    # unique at degenerate vertices, so they are not compared entry by entry)
    dropped = (np.abs(dual_on[:, :40 * 5 * 6]).reshape(len(agents), 40, 30).sum(axis=2) == 0)
    assert dropped.mean() > 0.5
    lsc_off = np.abs(dual_off[:, :40 * 5 * 6]).reshape(len(agents), 40, 30)
    assert lsc_off[dropped].max() < 1e-6
    _check_against_oracle(batch, agents[:8], off, normals, rhs, c_on, cost_on, st_on, min_checked=5)


def test_solve_parity_synthetic_planes():
    """config 4 shape: random half-spaces with a strictly feasible point"""
    batch = W.make_forest_batch(256, K=40, seed=20260004)
    off, normals, rhs = W.make_synthetic_planes(batch, K=40)
    agents = list(range(0, 256, 16))
    sel_off = np.array([0] + list(np.cumsum([off[a + 1] - off[a] for a in agents])), np.int32)
    sel_n = np.concatenate([normals[off[a]:off[a + 1]] for a in agents]); sel_r = np.concatenate([rhs[off[a]:off[a + 1]] for a in agents])
    planner = _planner(batch.cfg)
    ctrl, cost, status, iters, kkt, _ = _solve_host(planner, batch, agents, sel_off, sel_n, sel_r)
    _check_against_oracle(batch, agents, sel_off, sel_n, sel_r, ctrl, cost, status, min_checked=12)


def test_active_set_large_instance_takes_dense_neighbourhoods():
    """11..20 kept obstacles per agent (the dense spots of a closed loop): beyond the throughput active-set instance's
    capacity, solved by its large instance (row constants in shared memory), never by silently dropping rows; beyond 20
    the interior point takes over.  Both against the oracle."""
    batch = W.make_forest_batch(128, K=40, seed=20260004)
    for K, by_active_set in ((16, True), (28, False)):
        off, normals, rhs = W.make_synthetic_planes(batch, K=K)
        n = batch.n_agents
        planner = _planner(batch.cfg)
        agents = list(range(n))
        ctrl, cost, status, iters, kkt, _ = _solve_host(planner, batch, agents, off, normals, rhs)
        klass = planner.qp.last_instances(n)
        assert (status == 0).all()
        if by_active_set:
            assert (klass == 0).mean() > 0.9, np.bincount(klass)       # (a few may exceed 32 active rows: interior point)
        else:
            assert (klass == 2).mean() > 0.5, np.bincount(klass)       # reason 2: more kept obstacles than the instance holds
        sub = list(range(0, n, 8))
        sel_off = np.array([0] + list(np.cumsum([off[a + 1] - off[a] for a in sub])), np.int32)
        sel_n = np.concatenate([normals[off[a]:off[a + 1]] for a in sub]); sel_r = np.concatenate([rhs[off[a]:off[a + 1]] for a in sub])
        _check_against_oracle(batch, sub, sel_off, sel_n, sel_r, ctrl[sub], cost[sub], status[sub], min_checked=12)


def test_active_set_hand_over_on_hardware():
    """tests/golden/many_active_rows_case.npz (35 active rows at the optimum) replicated 1400 times: above the one-wave
    regime, so the throughput active-set instance runs first, stops at its 32-row capacity and hands its state over; the
    pool has 64 slots, so 64 copies are resumed by the large instance and the rest restarted by it -- every copy must
    land on the oracle's optimum, solved by the active-set passes"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "many_active_rows_case.npz"))
    n = 1400
    cfg = W.make_forest_batch(8, K=4).cfg
    cfg.world_min = (-66.0, -66.0, 0.0); cfg.world_max = (66.0, 66.0, 2.5)
    K = g["normals"].shape[0]
    qp = capi.LscQp(cfg, device=0)
    rep = lambda a: np.ascontiguousarray(np.repeat(a[None], n, axis=0))
    off = (np.arange(n + 1) * K).astype(np.int32)
    normals = np.ascontiguousarray(np.tile(g["normals"], (n, 1, 1))); rhs = np.ascontiguousarray(np.tile(g["rhs"], (n, 1, 1)))
    nv = cfg.dim * cfg.M * 6
    ctrl = np.zeros((n, nv)); cost = np.zeros(n); status = np.zeros(n, np.int32); iters = np.zeros(n, np.int32); kkt = np.zeros((n, 4))
    qp.solve_host(n, rep(g["state"]), rep(g["goal"]), rep(g["limits"]), None, off, normals, rhs, ctrl, cost, status, iters=iters, kkt=kkt)
    klass = qp.last_instances(n)
    assert (status == 0).all() and (klass == 0).all(), (np.bincount(status), np.bincount(klass))
    assert np.abs(ctrl - g["x"][None]).max() < 1e-8
    assert (kkt[:, 2].astype(int) // 64 > 32).all()           # every copy went beyond the throughput instance's 32 rows
    assert len(np.unique(iters)) == 1                          # resumed and restarted copies count the same iterations


def test_solve_with_sfc_boxes():
    """SFC rows (world_use_octomap): per-segment boxes around the previous solution"""
    cfg = W.PlannerConfig(use_sfc=True)
    batch = W.make_forest_batch(64, K=16, cfg=cfg)
    lo = batch.own_traj.min(axis=2) - 0.6; hi = batch.own_traj.max(axis=2) + 0.6
    batch.sfc = np.ascontiguousarray(np.concatenate([lo, hi], axis=-1).astype(np.float32))
    agents = list(range(0, 64, 8))
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    planner = _planner(batch.cfg)
    sfc = np.ascontiguousarray(batch.sfc[agents])
    ctrl, cost, status, iters, kkt, _ = _solve_host(planner, batch, agents, off, normals, rhs, sfc=sfc)
    _check_against_oracle(batch, agents, off, normals, rhs, ctrl, cost, status, sfc=sfc, min_checked=6)


def test_empty_and_ragged_obstacle_lists():
    """K = 0 for some agents, different K per agent, zero agents"""
    batch = W.make_forest_batch(64, K=40)
    agents = [0, 1, 2, 3, 4, 5]
    ks = [0, 1, 40, 7, 0, 23]
    cfgo = oracle_config(batch.cfg)
    normals, rhs, off = [], [], [0]
    for a, k in zip(agents, ks):
        pt, nr, d = oracle_lsc(batch, a, orc.GEN_LSC)
        n_, r_ = orc.pack_planes(cfgo, pt[:k], nr[:k], d[:k]) if k else (np.zeros((0, 5, 3)), np.zeros((0, 5, 6)))
        normals.append(n_); rhs.append(r_); off.append(off[-1] + k)
    normals = np.ascontiguousarray(np.concatenate(normals)); rhs = np.ascontiguousarray(np.concatenate(rhs)); off = np.array(off, np.int32)
    planner = _planner(batch.cfg)
    ctrl, cost, status, iters, kkt, _ = _solve_host(planner, batch, agents, off, normals, rhs)
    _check_against_oracle(batch, agents, off, normals, rhs, ctrl, cost, status, min_checked=5)
    # zero agents: a no-op that must not fail
    planner.qp.solve_host(0, batch.state[:0].copy(), batch.goal[:0].copy(), batch.limits[:0].copy(), None,
                          np.zeros(1, np.int32), None, None, np.zeros((0, 90)), np.zeros(0), np.zeros(0, np.int32))


def test_zero_normal_rows_are_skipped():
    """rows whose normal is shorter than SP_EPSILON_FLOAT are dropped (traj_optimizer.cpp:409-411)"""
    batch = W.make_forest_batch(64, K=8)
    agents = [3, 9]
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    normals[1] = 0.0; rhs[1] = 5.0            # would be infeasible if it were not skipped
    normals[off[1] + 2, 3] = 1e-7
    planner = _planner(batch.cfg)
    ctrl, cost, status, iters, kkt, _ = _solve_host(planner, batch, agents, off, normals, rhs)
    _check_against_oracle(batch, agents, off, normals, rhs, ctrl, cost, status, min_checked=2)


def test_infeasible_is_reported_not_nan():
    """contradictory half-spaces: status != OK, finite outputs (caller falls back to initial_traj, traj_planner.cpp:767-797)"""
    batch = W.make_forest_batch(64, K=2)
    agents = [0, 1]
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    normals[0, :, :] = [1.0, 0.0, 0.0]; rhs[0] = 100.0        # x >= 100
    normals[1, :, :] = [-1.0, 0.0, 0.0]; rhs[1] = 100.0       # -x >= 100
    planner = _planner(batch.cfg)
    ctrl, cost, status, iters, kkt, _ = _solve_host(planner, batch, agents, off, normals, rhs)
    assert status[0] != capi.STATUS_OK
    assert status[1] == capi.STATUS_OK
    assert np.isfinite(ctrl[1]).all()


@pytest.mark.parametrize("generator,M,dim", [(capi.GEN_LSC, 5, 3), (capi.GEN_CLSC, 10, 2), (capi.GEN_CLSC, 5, 3), (capi.GEN_BVC, 5, 3),
                                             (capi.GEN_RSFC, 5, 3)])
def test_assembly_parity(generator, M, dim):
    """device LSC assembly vs the oracle's restatement of generateLSC / generateCLSC / generateBVC"""
    import torch
    cfg = W.PlannerConfig(M=M, dim=dim)
    batch = W.make_forest_batch(96, K=24, cfg=cfg)
    if generator == capi.GEN_CLSC:
        near_goals(batch)
    planner = _planner(batch.cfg)
    d = planner.upload(batch)
    planner.assemble_device(d, generator)
    torch.cuda.synchronize()
    normals = d.normals.cpu().numpy(); rhs = d.rhs.cpu().numpy()
    agents = list(range(0, 96, 6))
    off, n_ref, r_ref = oracle_planes(batch, agents, generator)
    exact = 0; total = 0
    for i, a in enumerate(agents):
        sl = slice(batch.obs_offsets[a], batch.obs_offsets[a + 1])
        got_n, got_r = normals[sl], rhs[sl]
        want_n, want_r = n_ref[off[i]:off[i + 1]], r_ref[off[i]:off[i + 1]]
        # normals are float32 values: equal up to one float ulp of a unit vector
        assert np.abs(got_n - want_n).max() <= 2.4e-7, (a, np.abs(got_n - want_n).max())
        assert np.abs(got_r - want_r).max() <= 2e-6, (a, np.abs(got_r - want_r).max())
        exact += int((got_n == want_n).sum()); total += got_n.size
    assert exact / total > 0.99


def test_replan_host_matches_device_path_and_oracle():
    """fused host entry point: trajectories in, solutions out"""
    import torch
    batch = W.make_forest_batch(128, K=40)
    planner = _planner(batch.cfg)
    out = planner.replan_host(batch, capi.GEN_LSC)
    d = planner.upload(batch)
    planner.replan_device(d, capi.GEN_LSC)
    torch.cuda.synchronize()
    assert np.array_equal(out["ctrl"], d.ctrl.cpu().numpy())
    assert (out["status"] == 0).all()
    agents = [0, 17, 99]
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    _check_against_oracle(batch, agents, off, normals, rhs, out["ctrl"][agents], out["cost"][agents], out["status"][agents], min_checked=2)


def test_step_kernel_parity():
    """closed-loop glue: float narrowing, getStateAt, previous-solution shift"""
    import torch
    for M, dim in ((5, 3), (10, 2)):
        cfg = W.PlannerConfig(M=M, dim=dim)
        batch = W.make_forest_batch(64, K=4, cfg=cfg)
        planner = _planner(batch.cfg)
        cfgo = oracle_config(batch.cfg)
        rng = np.random.default_rng(1)
        traj = batch.own_traj.astype(np.float64) + rng.normal(0, 1e-3, batch.own_traj.shape)
        ctrl = np.ascontiguousarray(np.transpose(traj, (0, 3, 1, 2))[:, :dim].reshape(64, -1))
        dev = torch.device("cuda", 0)
        t_ctrl = torch.from_numpy(ctrl).to(dev)
        t_traj = torch.empty((64, M, 6, 3), dtype=torch.float32, device=dev)
        t_state = torch.empty((64, 9), dtype=torch.float32, device=dev)
        t_shift = torch.empty((64, M, 6, 3), dtype=torch.float32, device=dev)
        for step in (0.1, 0.2):
            planner.qp.step_batch(64, t_ctrl, step, t_traj, t_state, t_shift)
            torch.cuda.synchronize()
            got_traj, got_state, got_shift = t_traj.cpu().numpy(), t_state.cpu().numpy(), t_shift.cpu().numpy()
            for a in range(0, 64, 7):
                want = traj[a].astype(np.float32)
                if dim == 2:
                    want[..., 2] = np.float32(cfg.z_2d)
                assert np.array_equal(got_traj[a], want)
                st = orc.get_state_at(cfgo, want, step)
                if dim == 2:
                    st[2] = np.float32(cfg.z_2d)
                assert np.allclose(got_state[a], st, rtol=2e-6, atol=2e-6), (a, got_state[a], st)
                assert np.array_equal(got_shift[a], orc.shift_traj(cfgo, want))


def test_full_size_properties():
    """BASELINE size (4096 agents, K=40): size-independent properties of the solutions"""
    import torch
    batch = W.make_forest_batch(4096, K=40)
    planner = _planner(batch.cfg)
    d = planner.upload(batch)
    planner.assemble_device(d, capi.GEN_LSC)
    planner.solve_device(d, want_kkt=True)
    torch.cuda.synchronize()
    status = d.status.cpu().numpy(); ctrl = d.ctrl.cpu().numpy().reshape(4096, 3, 5, 6); kkt = d.kkt.cpu().numpy()
    assert (status == 0).all(), np.bincount(status)
    # equalities: initial state, C0/C1/C2 continuity, terminal stop
    st = batch.state.astype(np.float64); dt = batch.cfg.dt
    assert np.abs(ctrl[:, :, 0, 0] - st[:, 0:3]).max() < 1e-12
    assert np.abs(5 / dt * (ctrl[:, :, 0, 1] - ctrl[:, :, 0, 0]) - st[:, 3:6]).max() < 1e-9
    assert np.abs(20 / dt ** 2 * (ctrl[:, :, 0, 2] - 2 * ctrl[:, :, 0, 1] + ctrl[:, :, 0, 0]) - st[:, 6:9]).max() < 1e-7
    assert np.abs(ctrl[:, :, 1:, 0] - ctrl[:, :, :-1, 5]).max() < 1e-12
    assert np.abs((ctrl[:, :, 1:, 1] - ctrl[:, :, 1:, 0]) - (ctrl[:, :, :-1, 5] - ctrl[:, :, :-1, 4])).max() < 1e-12
    assert np.abs(ctrl[:, :, -1, 5] - ctrl[:, :, -1, 3]).max() == 0
    # inequalities: LSC rows, velocity / acceleration limits
    normals = d.normals.cpu().numpy().reshape(4096, 40, 5, 3); rhs = d.rhs.cpu().numpy().reshape(4096, 40, 5, 6)
    lhs = np.einsum("akmd,admi->akmi", normals, ctrl)
    viol = (rhs - lhs); viol[:, :, 0, :3] = -1
    assert viol.max() < 1e-8, viol.max()
    vel = 5 / dt * np.diff(ctrl, axis=3); acc = 20 / dt ** 2 * np.diff(ctrl, n=2, axis=3)
    assert np.abs(vel[:, :, 1:]).max() < 1 + 1e-7 and np.abs(acc[:, :, 1:]).max() < 2 + 1e-6
    # interior-point certificate from the kernel itself
    assert kkt[:, 1].max() < 1e-9 and kkt[:, 3].max() < 1e-10
    # idempotence: solving the same batch again is bit-identical (no atomics / order dependence)
    c1 = d.ctrl.clone()
    planner.solve_device(d)
    torch.cuda.synchronize()
    assert torch.equal(c1, d.ctrl)


def test_closed_loop_circle_swap_is_collision_free():
    """16 agents on a circle swapping positions (missions/*: antipodal goals), 60 replans of 0.2 s:
    no collision (safety ratio >= 1, what the reference's summary CSV reports) and progress towards the goals"""
    from lsc_dr_planner_b200.closed_loop import ClosedLoopSim
    n = 16
    cfg = W.PlannerConfig()
    batch = W.make_forest_batch(n, K=15, cfg=cfg, moving=False)
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    pos = np.stack([4 * np.cos(ang), 4 * np.sin(ang), np.full(n, 1.0)], 1).astype(np.float32)
    batch.state[:] = 0; batch.state[:, :3] = pos
    batch.goal = (-pos * [1, 1, -1]).astype(np.float32)
    batch.cfg.world_min = (-6.0, -6.0, 0.0); batch.cfg.world_max = (6.0, 6.0, 2.5)
    sim = ClosedLoopSim(batch, device=0, K=15)
    d0 = sim.max_goal_distance()
    worst = np.inf
    for _ in range(60):
        sim.step()
        worst = min(worst, sim.min_separation_ratio())
    assert worst >= 1.0 - 1e-3, worst
    assert sim.failed_total == 0
    assert sim.max_goal_distance() < 0.75 * d0


def test_cpp_shim_end_to_end():
    """the C++ class surfaces (include/lscqp_shim.hpp) driven like traj_planner.cpp drives the reference's classes"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "shim", "shim_smoke")
    libdir = os.path.join(root, "lsc_dr_planner_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(root, "tests", "shim", "shim_smoke.cpp"), "-o", exe,
                    "-L" + libdir, "-l:liblscqp.so", "-Wl,-rpath," + libdir, "-Wl,--allow-shlib-undefined"], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.splitlines()
    cps = np.array([[float(v) for v in l.split()[3:]] for l in out if l.startswith("cp ")]).reshape(5, 6, 3)
    assert "qpfailed 1" in out
    # the same QP through the oracle
    cfgo = orc.Config(world_min=(-5, -5, 0), world_max=(5, 5, 2.5))
    ag = orc.Agent(np.array([0, 0, 1]), np.zeros(3), np.zeros(3), np.array([3, 0, 1]))
    pt = np.tile(np.array([1, 0, 1], np.float32), (1, 5, 6, 1)); nr = np.tile(np.array([-1, 0, 0], np.float32), (1, 5, 6, 1))
    qp = orc.qp_build(cfgo, ag, pt, nr, np.full((1, 5, 6), 0.4))
    xe, ok = oracle_solution(qp)        # (polish may be refused on this degenerate toy model: raw HiGHS is ~1e-6 accurate)
    want = xe.reshape(3, 5, 6).transpose(1, 2, 0)
    assert np.abs(cps - want).max() < (1e-5 if ok else 1e-4)
    assert cps[..., 0].max() <= 0.6 + 1e-6
    cost = float([l for l in out if l.startswith("cost")][0].split()[1])
    assert abs(cost - (xe @ qp.P @ xe + qp.q @ xe + qp.c0)) < 1e-6 * max(1.0, cost)
    g = [float(v) for v in [l for l in out if l.startswith("goal ")][0].split()[1:]]
    # t* = 0.5 on the segment from the waypoint (1, 0, 1) to the previous goal (0.2, 0.5, 1)
    assert np.abs(np.array(g) - np.array([0.6, 0.25, 1.0])).max() < 1e-6 and "goalfailed 1" in out
    b = [l for l in out if l.startswith("batch")][0].split()
    assert b[1] == "0" and b[2] == "0" and float(b[3]) > 0.3 and float(b[4]) < 2.7
    # CollisionConstraints::initializeSFC / constructSFCFromPoint / constructSFCFromConvexHull against the oracle, bit for bit
    m = orc.Map(np.array([[2.0, 0.0, 1.25, 0.5, 0.5, 2.5]]), (-5, -5, 0), (5, 5, 2.5))
    line = lambda key: [l for l in out if l.startswith(key + " ")][0].split()[1:]
    f32 = lambda vals: np.array([float(v) for v in vals], np.float32)
    ok, box = m.sfc_initialize((1.0, 0.3, 1.0), 0.15)
    v = line("sfc_init")
    assert ok and np.array_equal(f32(v[:6]), box) and v[7] == "1"
    st1, box1 = m.sfc_from_point((1.2, 0.35, 1.0), (3.0, 2.0, 1.0), box, 0.15)
    v = line("sfc_point")
    assert int(v[0]) == st1 and np.array_equal(f32(v[1:]), box1)
    st2, box2 = m.sfc_from_convex_hull([(1.2, 0.35, 1.0), (1.3, 0.5, 1.0)], (1.4, 1.0, 1.0), box1, 0.15)
    v = line("sfc_hull")
    assert int(v[0]) == st2 and np.array_equal(f32(v[1:]), box2)
    assert "sfc_invalid 1" in out
    xq, xmax = (float(t) for t in line("sfc_qp"))
    okq, boxq = m.sfc_initialize((1.3, 0.0, 1.0), 0.15)
    cfgs = orc.Config(world_min=(-5, -5, 0), world_max=(5, 5, 2.5), use_sfc=True)
    ags = orc.Agent(np.array([1.3, 0, 1]), np.zeros(3), np.zeros(3), np.array([4, 0, 1]))
    qps = orc.qp_build(cfgs, ags, np.zeros((0, 5, 6, 3), np.float32), np.zeros((0, 5, 6, 3), np.float32), np.zeros((0, 5, 6)), np.tile(boxq, (5, 1)))
    xs, oks = oracle_solution(qps)
    # the corridor stops short of the pillar (x_max 1.65 < 1.75) and bounds the QP: without it the end point would pass 1.75
    assert okq and abs(xmax - boxq[3]) < 1e-7 and xmax < 1.7 and xq <= xmax + 1e-6 and abs(xq - xs.reshape(3, 5, 6)[0, -1, -1]) < 1e-5


@pytest.mark.parametrize("M,dim,K,rng_,mode", [(10, 2, 9, 0.7, 1), (5, 3, 12, 0.7, 1), (10, 2, 9, 3.0, 1),
                                               (5, 3, 12, 3.0, 0), (5, 3, 12, 0.7, 0), (10, 2, 9, 0.7, 2)])
def test_communication_range_rows(M, dim, K, rng_, mode):
    """the launch-file shape (M=10, 2-D, communication_range 3, generateCLSC) and a binding range: rows of
    traj_optimizer.cpp:477-500 through the dense instance, assembly on the device.  mode 0 / range 3 is the reference's
    own default parameter set (param.cpp:117,129: mode/planner = dlsc, communication/range = 3.0)."""
    import torch
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=mode, comm_range=rng_)
    batch = W.make_forest_batch(64, K=K, cfg=cfg)
    r = np.random.default_rng(11)
    d = r.normal(size=(64, 3)); d[:, 2] = 0; d /= np.linalg.norm(d, axis=1, keepdims=True)
    batch.goal = (batch.state[:, :3] + (5.0 if rng_ < 1 else 0.6) * d).astype(np.float32)
    batch.next_waypoint = (batch.state[:, :3] + r.uniform(-0.05, 0.05, (64, 3))).astype(np.float32)
    if dim == 2:
        batch.next_waypoint[:, 2] = cfg.z_2d; batch.goal[:, 2] = cfg.z_2d
    gen = {1: capi.GEN_CLSC, 0: capi.GEN_LSC, 2: capi.GEN_BVC}[mode]      # constructLSC, traj_planner.cpp:552-569
    planner = _planner(batch.cfg)
    dev = planner.upload(batch)
    planner.replan_device(dev, gen)
    torch.cuda.synchronize()
    status = dev.status.cpu().numpy(); ctrl = dev.ctrl.cpu().numpy(); cost = dev.cost.cpu().numpy()
    # a 0.2 m leash can be infeasible for agents that cannot brake in time: those must be *reported*, the rest solved
    ok_agents = np.where(status == 0)[0]
    assert len(ok_agents) >= (64 if rng_ > 1 else 8), np.bincount(status)
    agents = [int(a) for a in ok_agents[:4]]
    off, normals, rhs = oracle_planes(batch, agents, gen)
    _check_against_oracle(batch, agents, off, normals, rhs, ctrl[agents], cost[agents], status[agents], min_checked=2)
    for a in np.where(status != 0)[0][:2]:         # the oracle agrees that the reported ones have no solution
        o2, n2, r2 = oracle_planes(batch, [int(a)], gen)
        assert orc.solve_highs(oracle_qp_from_planes(batch, int(a), n2, r2)).status != "Optimal"
    if rng_ < 1:      # the range must actually bind: end points stay within range/2 - radius of the start
        end = ctrl.reshape(64, dim, M, 6)[ok_agents][:, :, :, 5]
        dist = np.abs(end - batch.state[ok_agents][:, :dim, None]).max(axis=(1, 2))
        assert dist.max() <= 0.5 * rng_ - 0.15 + 1e-7 and dist.max() > 0.5 * rng_ - 0.15 - 1e-4


def test_replan_host_rejects_bad_neighbour_lists():
    """lscqp_replan_host validates the CSR neighbour lists before anything is launched (nothing written)"""
    from lsc_dr_planner_b200.planner import BatchPlanner
    batch = W.make_forest_batch(16, K=4)
    planner = BatchPlanner(batch.cfg, device=0)
    b = planner.host_buffers(batch)
    b["obs_index"][3] = 16                                    # outside [0, n_agents)
    with pytest.raises(capi.LscqpError, match="obs_index"):
        planner.replan_host_buffers(b, 16)
    b["obs_index"][3] = 0
    b["obs_offsets"][1] = 60                                  # 60 obstacles for agent 0 > max_obs
    with pytest.raises(capi.LscqpError, match="max_obs"):
        planner.replan_host_buffers(b, 16)


def test_solve_and_goal_host_reject_lists_above_capacity_and_device_path_reports_them():
    """no entry point shortens an obstacle list: lscqp_solve_host / lscqp_goal_host return LSCQP_E_CAPACITY (the shim then
    throws QPFAILED), the device entry point reports the agent through status_out = LSCQP_CAPACITY"""
    import torch
    cfg = W.PlannerConfig(max_obs=12)
    batch = W.make_forest_batch(64, K=12, cfg=cfg)
    agents = [0, 1, 2]
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    k = 12
    n1 = np.concatenate([normals[off[1]:off[2]], normals[off[1]:off[1] + 1]]); r1 = np.concatenate([rhs[off[1]:off[2]], rhs[off[1]:off[1] + 1]])
    normals2 = np.ascontiguousarray(np.concatenate([normals[:off[1]], n1, normals[off[2]:]]))
    rhs2 = np.ascontiguousarray(np.concatenate([rhs[:off[1]], r1, rhs[off[2]:]]))
    off2 = np.array([0, k, 2 * k + 1, 3 * k + 1], np.int32)
    planner = _planner(cfg)
    with pytest.raises(capi.LscqpError, match="max_obs"):
        _solve_host(planner, batch, agents, off2, normals2, rhs2)
    st = np.ascontiguousarray(batch.state[agents]); goal = np.ascontiguousarray(batch.goal[agents])
    gout = np.zeros((3, 3), np.float32); gst = np.zeros(3, np.int32)
    with pytest.raises(capi.LscqpError, match="max_obs"):
        planner.qp.goal_host(3, goal, goal, None, off2, normals2, rhs2, gout, gst)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ctrl = torch.zeros((3, 90), dtype=torch.float64, device="cuda"); cost = torch.zeros(3, dtype=torch.float64, device="cuda")
    status = torch.full((3,), -1, dtype=torch.int32, device="cuda")
    warm = t(batch.own_traj[agents])
    planner.qp.solve_batch(3, t(st), t(goal), t(batch.limits[agents]), None, t(off2), t(normals2), t(rhs2), ctrl, cost, status,
                           initial_traj=warm)
    torch.cuda.synchronize()
    assert status.cpu().tolist() == [0, capi.STATUS_CAPACITY, 0]
    ref = _solve_host(planner, batch, agents, off, normals, rhs, warm=True)
    got = ctrl.cpu().numpy()
    assert np.array_equal(got[0], ref[0][0]) and np.array_equal(got[2], ref[0][2]) and np.isfinite(got[1]).all()


@pytest.mark.parametrize("n_total,lo,n_local,K,comm", [(4096, 0, 4096, 40, 0.0), (1024, 256, 512, 40, 3.0), (15000, 14000, 1000, 17, 0.0),
                                                       (4096, 0, 4096, 40, 6.0), (3000, 100, 2000, 5, 4.0)])
def test_neighbour_selection_gpu(n_total, lo, n_local, K, comm):
    """lscqp_select_neighbours against a numpy restatement of the broadcastMsgs filter: ragged CSR lists holding exactly
    the in-range agents; K nearest + overflow flag only where more than K are in range"""
    import torch
    from test_emul_kernels import _check_neighbours
    rng = np.random.default_rng(n_total)
    state = np.zeros((n_total, 9), np.float32)
    state[:, :3] = rng.uniform(-40, 40, (n_total, 3)).astype(np.float32)
    cfg = W.PlannerConfig()
    qp = capi.LscQp(cfg, device=0)
    st = torch.from_numpy(state).cuda()

    def run():
        off = torch.zeros((n_local + 1,), dtype=torch.int32, device="cuda")
        idx = torch.full((n_local * K,), -1, dtype=torch.int32, device="cuda")
        over = torch.zeros((n_local,), dtype=torch.int32, device="cuda")
        qp.select_neighbours(n_total, lo, n_local, K, comm, st, off, idx, over)
        torch.cuda.synchronize()
        return off.cpu().numpy(), idx.cpu().numpy(), over.cpu().numpy()
    off, idx, over = run()
    sel = range(n_local) if n_local <= 512 else [int(r) for r in rng.choice(n_local, 64, replace=False)]
    _check_neighbours(state, lo, off, idx[:off[-1]], over, K, comm, rows=sel)
    assert (np.diff(off) <= K).all() and (idx[off[-1]:] == -1).all()
    off2, idx2, over2 = run()
    assert np.array_equal(off, off2) and np.array_equal(idx, idx2) and np.array_equal(over, over2)    # deterministic


def test_closed_loop_cuda_graph_matches_eager():
    """the captured step (library launches + torch glue in one CUDA graph) reproduces the eager loop bit for bit"""
    import torch
    from lsc_dr_planner_b200.closed_loop import ClosedLoopSim
    sims = []
    for graph, exchange in ((False, "p2p"), (True, "p2p"), (False, "nccl")):
        batch = W.make_forest_batch(96, K=20, moving=False, seed=7)
        batch.goal = (batch.state[:, :3] * [-1, -1, 1]).astype(np.float32)
        sim = ClosedLoopSim(batch, device=0, K=20, use_graph=graph, exchange=exchange)
        for _ in range(12):
            sim.step()
        sim.sync_state()
        sims.append(sim)
    assert sims[1]._graph is not None and sims[0].exchange == "p2p" and sims[2].exchange == "nccl"
    assert torch.equal(sims[0].state, sims[1].state) and torch.equal(sims[0].traj, sims[1].traj)
    # the fused failsafe + step + publish kernel against the separate torch failsafe + lscqp_step_batch path
    assert torch.equal(sims[0].state, sims[2].state) and torch.equal(sims[0].traj, sims[2].traj)
    assert sims[0].failed_total == sims[1].failed_total == sims[2].failed_total == 0
    assert sims[0].exchange_timeouts == 0 and float((sims[0].state[:, :3] - torch.from_numpy(batch.state[:, :3]).cuda()).abs().max()) > 0.5


@pytest.mark.parametrize("generator,M,dim", [(capi.GEN_LSC, 5, 3), (capi.GEN_CLSC, 5, 3), (capi.GEN_CLSC, 10, 2)])
def test_fused_pruned_assembly_is_exact_gpu(generator, M, dim):
    """the replan path's assembly (neighbours read in place, provably inactive pairs dropped) against the materialised
    one: kept planes bit-identical, dropped pairs zero, and the QP solutions unchanged at full batch size"""
    import torch
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=capi.MODE_LSC)
    batch = W.make_forest_batch(1024, K=40 if M == 5 else 9, cfg=cfg)
    near_goals(batch)
    planner = _planner(batch.cfg)
    d = planner.upload(batch)
    planner.assemble_device(d, generator)
    planner.solve_device(d)
    torch.cuda.synchronize()
    n_ref, r_ref, c_ref, s_ref = d.normals.clone(), d.rhs.clone(), d.ctrl.clone(), d.status.clone()
    planner.assemble_fused_device(d, generator, prune=False)
    torch.cuda.synchronize()
    assert torch.equal(d.normals, n_ref) and torch.equal(d.rhs, r_ref)
    planner.assemble_fused_device(d, generator, prune=True)
    planner.solve_device(d)
    torch.cuda.synchronize()
    dropped = (d.normals == 0).all(dim=2) & ~(n_ref == 0).all(dim=2)
    assert torch.equal(d.normals[~dropped], n_ref[~dropped]) and torch.equal(d.rhs[~dropped], r_ref[~dropped])
    assert float(dropped.float().mean()) > (0.5 if dim == 3 else -1.0)
    ok = (s_ref == 0) & (d.status == 0)
    assert int(ok.sum()) >= int((s_ref == 0).sum())
    assert float((d.ctrl[ok] - c_ref[ok]).abs().max()) < 1e-6


def test_validate_batch_gpu():
    """lscqp_validate_batch (isSolValid) on the outputs of lscqp_step_batch, against the oracle"""
    import torch
    cfg = W.PlannerConfig(M=5, dim=3, planner_mode=capi.MODE_DLSC, use_sfc=True)
    batch = W.make_forest_batch(256, K=8, cfg=cfg)
    near_goals(batch)
    rng = np.random.default_rng(9)
    n, M = batch.n_agents, cfg.M
    lo = batch.own_traj.min(axis=2) - rng.uniform(0.05, 0.6, (n, M, 3)); hi = batch.own_traj.max(axis=2) + rng.uniform(0.05, 0.6, (n, M, 3))
    batch.sfc = np.ascontiguousarray(np.concatenate([lo, hi], axis=2).astype(np.float32))
    batch.limits[::4, :3] = 0.2                                   # tight velocity limits: some solutions end up not valid
    planner = _planner(cfg)
    d = planner.upload(batch)
    planner.replan_device(d)
    dev = d.ctrl.device
    traj = torch.empty((n, M, 6, 3), dtype=torch.float32, device=dev); state = torch.empty((n, 9), dtype=torch.float32, device=dev)
    valid = torch.zeros(n, dtype=torch.int32, device=dev)
    planner.qp.step_batch(n, d.ctrl, cfg.dt, traj, state)
    state[1::7, 3:6] *= 3.0
    planner.qp.validate_batch(n, traj, state, d.limits, d.sfc, valid)
    torch.cuda.synchronize()
    cfgo = oracle_config(cfg)
    tr, st, got = traj.cpu().numpy(), state.cpu().numpy(), valid.cpu().numpy()
    want = np.array([orc.is_sol_valid(cfgo, orc.Agent(st[a, :3], st[a, 3:6], st[a, 6:9], batch.goal[a], max_vel=tuple(batch.limits[a, :3]),
                                                      max_acc=tuple(batch.limits[a, 3:6])), tr[a], st[a], batch.sfc[a]) for a in range(n)], np.int32)
    assert np.array_equal(got, want) and 0 < want.sum() < n


def test_light_and_full_instances_agree_at_full_size():
    """BASELINE's 4096-agent batch: the interior-point two-pass dispatch (one-warp light instance first), the one-pass
    full-capacity instance and the default dispatch (dual active-set first pass) return the same solutions, statuses and
    (to the last iterations' noise) objective values"""
    import copy
    import torch
    batch = W.make_forest_batch(4096, K=40)
    outs = []
    for presolve in (9, 11, 1):                              # 9 = light + full, 11 = full only, 1 = default (active set first)
        cfg = copy.copy(batch.cfg); cfg.presolve = presolve
        planner = _planner(cfg)
        d = planner.upload(batch)
        planner.replan_device(d)
        torch.cuda.synchronize()
        outs.append((d.ctrl.clone(), d.status.clone(), d.cost.clone(), d.iters.clone()))
    (c1, s1, f1, i1), (c3, s3, f3, i3), (ca, sa, fa, ia) = outs
    assert int((s1 != 0).sum()) == 0 and int((s3 != 0).sum()) == 0 and int((sa != 0).sum()) == 0
    assert float((ca - c3).abs().max()) < 2e-6
    assert float(((fa - f3).abs() / f3.abs().clamp(min=1.0)).max()) < 1e-8
    assert float((c1 - c3).abs().max()) < 2e-6
    assert float(((f1 - f3).abs() / f3.abs().clamp(min=1.0)).max()) < 1e-8
    assert abs(float(i1.float().mean()) - float(i3.float().mean())) < 0.2


@pytest.mark.parametrize("solver", ["das", "pdip"])
def test_bench_batch_light_instance_and_pruned_assembly_against_oracle(solver):
    """The kernels the bench number is made of, against the oracle on hardware: BASELINE's 4096-agent batch through the
    fused pruned assembly and the two-pass solve; 64 sampled agents that the first pass solved -- the dual active-set
    kernel of the default dispatch ("das"), or the one-warp *light* interior-point instance with it switched off ("pdip")
    (lscqp_last_instances) are compared with the oracle's optimum of the reference's full model (all 40 obstacles, all
    rows) at 1e-5 m, and their pruned planes with the oracle's planes: a kept (obstacle, segment) pair carries the
    reference's normal (<= 1 float ulp) and constants, a dropped pair (zero normal) is strictly inactive at the optimum."""
    import torch
    batch = W.make_forest_batch(4096, K=40)
    if solver == "pdip":
        batch.cfg.presolve = PDIP_ONLY
    planner = _planner(batch.cfg)
    d = planner.upload(batch)
    planner.replan_device(d)
    torch.cuda.synchronize()
    klass = planner.qp.last_instances(4096)
    assert (klass == 0).mean() > 0.9                          # the first pass is what the bench measures
    status = d.status.cpu().numpy(); ctrl = d.ctrl.cpu().numpy(); cost = d.cost.cpu().numpy()
    normals_d = d.normals.cpu().numpy(); rhs_d = d.rhs.cpu().numpy()
    assert (status == 0).all()
    rng = np.random.default_rng(64)
    agents = [int(a) for a in rng.choice(np.where(klass == 0)[0], 64, replace=False)]
    agents += [int(a) for a in np.where(klass != 0)[0][:4]]   # and a few the full-capacity pass took over
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    errs = _check_against_oracle(batch, agents, off, normals, rhs, ctrl[agents], cost[agents], status[agents], min_checked=60)
    print("light-instance parity: max |ctrl - oracle| = %.2e, median %.2e over %d agents" % (errs.max(), np.median(errs), len(errs)))
    M = batch.cfg.M
    kept = dropped = 0
    for i, a in enumerate(agents[:16]):
        sl_o = slice(off[i], off[i + 1]); sl_d = slice(batch.obs_offsets[a], batch.obs_offsets[a + 1])
        nd, rd, no, ro = normals_d[sl_d], rhs_d[sl_d], normals[sl_o], rhs[sl_o]
        zero = (nd == 0).all(axis=2)                          # [K, M] pairs the pruning dropped
        assert np.abs(nd[~zero] - no[~zero]).max() <= 2.4e-7 and np.abs(rd[~zero] - ro[~zero]).max() <= 2e-6
        x = ctrl[a].reshape(3, M, 6)
        q = np.einsum("kmd,dmi->kmi", no, x) - ro             # row values of the reference's planes at the optimum
        q[:, 0, :3] = np.inf                                  # (no rows on the first three control points)
        assert q[zero].min() > 1e-6, (a, q[zero].min())
        kept += int((~zero).sum()); dropped += int(zero.sum())
    assert kept > 0 and dropped > kept


@pytest.mark.parametrize("which", [0, 1])
def test_map_and_sfc_kernels_gpu(which):
    """SURVEY row f2 on hardware: lscqp_map_set (occupancy + nearest-obstacle field of a reference world) and
    lscqp_sfc_batch (initializeSFC / constructSFCFromPoint / constructSFCFromConvexHull for 64 agents at once) against
    the oracle's restatement, bit for bit"""
    import torch
    from test_sfc import _agents_in_free_space, worlds
    name, boxes = worlds()[which]
    wmin, wmax = (-5.0, -5.0, 0.0), (5.0, 5.0, 2.5)
    cfg = W.PlannerConfig(M=5, dim=3, world_min=wmin, world_max=wmax, use_sfc=True)
    m = orc.Map(boxes, wmin, wmax)
    qp = capi.LscQp(cfg, device=0)
    qp.map_set(boxes, 0.1, 1.0)
    occ, closest = qp.map_get()
    assert np.array_equal(occ, m.occupancy())
    cl = m.closest()
    packed = np.where(cl[..., 0] < 0, -1, cl[..., 0] | (cl[..., 1] << 10) | (cl[..., 2] << 20)).astype(np.int32)
    assert np.array_equal(closest, packed)
    rng = np.random.default_rng(10 + which)
    n, radius, M = 64, 0.15, cfg.M
    pos = _agents_in_free_space(m, rng, n - 1, wmin, wmax, radius)
    pos = np.concatenate([pos, np.array([[boxes[0, 0], boxes[0, 1], 1.0]], np.float32)])      # the last agent sits inside a box
    lim = np.tile(np.array([1, 1, 1, 2, 2, 2, radius, 1.0]), (n, 1))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_lim = t(lim)
    sfc = torch.full((n, M, 6), 7.0, dtype=torch.float32, device="cuda"); status = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    qp.sfc_batch(capi.SFC_INIT, n, t(pos), None, None, d_lim, sfc, status)
    torch.cuda.synchronize()
    got, st = sfc.cpu().numpy(), status.cpu().numpy()
    for a in range(n):
        ok, box = m.sfc_initialize(pos[a], radius)
        assert st[a] == int(ok)
        assert all(np.array_equal(got[a, s_], box) for s_ in range(M)) if ok else (got[a] == 7.0).all()
    assert st[-1] == 0 and st[:-1].all()
    n = n - 1; pos = pos[:n]; lim = lim[:n]; d_lim = t(lim)
    for mode in (capi.SFC_FROM_POINT, capi.SFC_FROM_HULL):
        cur = got[:n].copy()
        for step in range(3):
            last = (pos + rng.uniform(-0.5, 0.5, pos.shape)).astype(np.float32)
            goal = (last + rng.uniform(-1.2, 1.2, pos.shape)).astype(np.float32)
            wp = (goal + rng.uniform(-0.3, 0.3, pos.shape)).astype(np.float32)
            for arr in (last, goal, wp):
                arr[:, 2] = np.clip(arr[:, 2], 0.3, 2.2)
            want = cur.copy(); want_st = np.zeros(n, np.int32)
            for a in range(n):
                prev = cur[a, M - 1].copy()
                if mode == capi.SFC_FROM_POINT:
                    s_, box = m.sfc_from_point(last[a], goal[a], prev, radius)
                else:
                    s_, box = m.sfc_from_convex_hull([last[a], goal[a]], wp[a], prev, radius)
                want[a, :M - 1] = cur[a, 1:]; want[a, M - 1] = box; want_st[a] = s_
            d_sfc = t(cur); d_st = torch.full((n,), -1, dtype=torch.int32, device="cuda")
            qp.sfc_batch(mode, n, t(last), t(goal), t(wp), d_lim, d_sfc, d_st)
            torch.cuda.synchronize()
            assert np.array_equal(d_st.cpu().numpy(), want_st), (mode, step)
            cur = d_sfc.cpu().numpy()
            assert np.array_equal(cur, want), (mode, step)
        assert (want_st > 0).mean() > 0.5                      # mostly fresh corridors, not the reuse-previous fallback


def test_closed_loop_with_safe_flight_corridors_avoids_the_forest():
    """forest10 mission in the reference's forest1 world with world_use_octomap on: corridors built on the device every
    replan (lscqp_sfc_batch), right-hand-rule goals; no agent-agent collision, every control point inside its corridor, no
    agent ever inside an occupied cell, no QP failure, the agents make progress.
    (Clearance: isObstacleInSFC tests only the Euclidean-NEAREST occupied cell of a grid point's cell in the L-inf metric
    (collision_constraints.cpp:795-803), so a corridor corner can end 0.05 m from the corner cell of a pillar whose other
    cells are farther by centre distance -- restated as is; the reference never evaluates safety_ratio_obs, its summary
    logs print 1e+09.  Observed minimum here: 0.078 m.)"""
    from test_missions import FOREST10
    from test_sfc import worlds
    from lsc_dr_planner_b200 import missions as MS
    from lsc_dr_planner_b200.closed_loop import ClosedLoopSim
    boxes = worlds()[0][1]
    mission = MS.parse_mission(FOREST10, 3, 1.0)
    cfg = MS.launch_config(mission, M=5, dim=3, comm_range=0.0)
    cfg.use_sfc = True
    batch = MS.first_replan_batch(mission, cfg)
    batch.goal = mission.goal.copy()
    sim = ClosedLoopSim(batch, device=0, K=9, goal_mode="righthand", world_boxes=boxes)
    m = orc.Map(boxes, cfg.world_min, cfg.world_max)
    occ = np.argwhere(m.occupancy())
    lo = (occ + np.array(m.key0)) * 0.1; hi = lo + 0.1
    d0 = sim.max_goal_distance()
    worst_agents, worst_obs = np.inf, np.inf
    for _ in range(80):
        sim.step()
        worst_agents = min(worst_agents, sim.min_separation_ratio())
        box = sim.sfc.cpu().numpy(); traj = sim.traj_out.cpu().numpy()            # this replan's corridors and solutions
        inside = (traj >= box[:, :, None, :3] - 2e-5) & (traj <= box[:, :, None, 3:] + 2e-5)
        inside[:, 0, :3] = True                                                    # (fixed by the initial state, traj_optimizer.cpp:260-263)
        assert inside.all()
        p = sim.state[:, :3].cpu().numpy().astype(np.float64)
        dist = np.sqrt((np.maximum(np.maximum(lo[None] - p[:, None], p[:, None] - hi[None]), 0) ** 2).sum(-1)).min()
        worst_obs = min(worst_obs, dist)
    assert int(sim._sfc_invalid.item()) == 0 and sim.failed_total == 0
    assert worst_agents >= 1.0 - 1e-3 and worst_obs > 0.04, (worst_agents, worst_obs)
    assert sim.max_goal_distance() < 0.75 * d0                                # (without a grid planner some agents queue behind pillars)


def test_infeasible_case_from_the_sweep_gpu():
    """the captured infeasible CLSC agent (tests/golden/infeasible_case.npz): reported, finite, stopped early"""
    from test_emul_kernels import _infeasible_case
    cfg, g, off = _infeasible_case()
    qp = capi.LscQp(cfg, device=0)
    nv = cfg.dim * cfg.M * 6
    ctrl = np.zeros((1, nv)); cost = np.zeros(1); status = np.zeros(1, np.int32); iters = np.zeros(1, np.int32); kkt = np.zeros((1, 4))
    qp.solve_host(1, g["state"][None].copy(), g["goal"][None].copy(), g["limits"][None].copy(), None, off,
                  np.ascontiguousarray(g["normals"]), np.ascontiguousarray(g["rhs"]), ctrl, cost, status, iters=iters, kkt=kkt,
                  initial_traj=g["own"][None].copy())
    assert status[0] in (2, 3) and iters[0] < 45
    assert np.isfinite(ctrl).all() and np.isfinite(cost).all() and np.isfinite(kkt).all()


def test_closed_loop_right_hand_rule_completes_the_forest10_swap():
    """the reference's forest10 mission (10 agents swapping across a circle) with the right-hand-rule goal mode
    (traj_planner.cpp:468-477): every agent reaches its goal, collision-free, no QP failure; with static goals the
    symmetric swap deadlocks (also collision-free)"""
    from test_missions import FOREST10
    from lsc_dr_planner_b200 import missions as MS
    from lsc_dr_planner_b200.closed_loop import ClosedLoopSim
    mission = MS.parse_mission(FOREST10, 3, 1.0)
    cfg = MS.launch_config(mission, M=5, dim=3, comm_range=0.0)
    results = {}
    for mode in ("righthand", "static"):
        batch = MS.first_replan_batch(mission, cfg)
        batch.goal = mission.goal.copy()
        sim = ClosedLoopSim(batch, device=0, K=9, goal_mode=mode)
        worst, steps = np.inf, 0
        for steps in range(1, 121):
            sim.step()
            worst = min(worst, sim.min_separation_ratio())
            if sim.max_goal_distance() < 0.1:
                break
        results[mode] = (steps, worst, sim.max_goal_distance(), sim.failed_total)
    steps, worst, dist, failed = results["righthand"]
    assert dist < 0.1 and steps < 100 and worst >= 1.0 - 1e-3 and failed == 0, results
    steps, worst, dist, failed = results["static"]
    assert dist > 1.0 and worst >= 1.0 - 1e-3 and failed == 0, results
