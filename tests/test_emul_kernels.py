"""CPU tests that step the *unmodified* CUDA kernel sources on the cooperative thread emulator
(tests/cuda_emul) and compare them with the oracle -- the same assertions the GPU parity tests make,
at sizes the emulator finishes in seconds."""
import os

import numpy as np
import pytest

import emul
from common import near_goals, oracle_config, oracle_planes, oracle_qp_from_planes, oracle_solution
from lsc_dr_planner_b200 import workloads as W
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,cfg,K", [("m5d3lsc", W.PlannerConfig(M=5, dim=3, planner_mode=1), 40),
                                        ("m10d2lsc", W.PlannerConfig(M=10, dim=2, planner_mode=1), 9),
                                        ("m5d3dlsc", W.PlannerConfig(M=5, dim=3, planner_mode=0), 16)])
def test_solve_kernel_matches_golden_solutions(name, cfg, K):
    g = np.load(os.path.join(GOLDEN, "qp_golden.npz"))
    batch = W.make_forest_batch(64, K=K, cfg=cfg)
    off, normals, rhs = g[name + "_off"], g[name + "_normals"], g[name + "_rhs"]
    agents = [0, 5, 11, 17, 23, 42]
    st = np.ascontiguousarray(batch.state[agents]); goal = np.ascontiguousarray(batch.goal[agents]); lim = np.ascontiguousarray(batch.limits[agents])
    sel = [int(i) for i in g[name + "_keep"][:3]]
    for warm in (False, True):
        it = np.ascontiguousarray(batch.own_traj[agents]) if warm else None
        ctrl, cost, status, iters, kkt, _ = emul.solve_batch(batch.cfg, len(agents), st, goal, lim, None,
                                                             np.ascontiguousarray(off), np.ascontiguousarray(normals),
                                                             np.ascontiguousarray(rhs), initial_traj=it)
        assert (status == 0).all()
        for x, i in zip(g[name + "_x"][:3], sel):
            assert np.abs(ctrl[i] - x).max() < 1e-5, (name, warm, i, np.abs(ctrl[i] - x).max())
        assert kkt[:, 1].max() < 1e-9 and kkt[:, 3].max() < 1e-10


def test_active_set_instances_by_kept_obstacles():
    """dispatch of the dual active-set kernel on the emulator: up to 10 kept obstacles the throughput instance, 11..20 the
    large instance (row constants in shared memory), beyond that the interior point -- the same optimum on every route"""
    from common import oracle_qp_from_planes, oracle_solution
    batch = W.make_forest_batch(64, K=40)
    agents = [0, 7, 12, 33]
    st = np.ascontiguousarray(batch.state[agents]); goal = np.ascontiguousarray(batch.goal[agents]); lim = np.ascontiguousarray(batch.limits[agents])
    for K, solved_by_das in ((8, 4), (16, 4), (30, 0)):
        off, nrm, rhs = W.make_synthetic_planes(batch, K=K)
        sel_off = np.array([0] + list(np.cumsum([off[a + 1] - off[a] for a in agents])), np.int32)
        sn = np.ascontiguousarray(np.concatenate([nrm[off[a]:off[a + 1]] for a in agents]))
        sr = np.ascontiguousarray(np.concatenate([rhs[off[a]:off[a + 1]] for a in agents]))
        ctrl, cost, status, iters, kkt, _ = emul.solve_batch(batch.cfg, len(agents), st, goal, lim, None, sel_off, sn, sr)
        assert (status == 0).all()
        assert emul.lib().emul_das_solved() == solved_by_das, (K, emul.lib().emul_das_solved())
        for i, a in enumerate(agents):
            xe, ok = oracle_solution(oracle_qp_from_planes(batch, a, sn[sel_off[i]:sel_off[i + 1]], sr[sel_off[i]:sel_off[i + 1]]))
            assert ok and np.abs(ctrl[i] - xe).max() < 1e-5, (K, a, np.abs(ctrl[i] - xe).max())


def test_active_set_many_active_rows_case():
    """tests/golden/many_active_rows_case.npz: the agent of the 8-GPU bench batches (seed 20260007, agent 2515) whose optimum
    has 35 active rows -- more than the throughput active-set instance holds (32): taken over by the large instance
    (two active rows per lane), not by the slower interior point; same optimum as the oracle's"""
    g = np.load(os.path.join(GOLDEN, "many_active_rows_case.npz"))
    cfg = W.make_forest_batch(8, K=4).cfg
    cfg.world_min = (-66.0, -66.0, 0.0); cfg.world_max = (66.0, 66.0, 2.5)
    off = np.array([0, g["normals"].shape[0]], np.int32)
    ctrl, cost, status, iters, kkt, _ = emul.solve_batch(cfg, 1, g["state"][None].copy(), g["goal"][None].copy(), g["limits"][None].copy(),
                                                         None, off, np.ascontiguousarray(g["normals"]), np.ascontiguousarray(g["rhs"]))
    assert status[0] == 0 and emul.lib().emul_das_solved() == 1
    assert int(kkt[0, 2]) // 64 > 32                                  # largest active set of the run
    assert np.abs(ctrl[0] - g["x"]).max() < 1e-8


def test_solve_kernel_duals_give_a_kkt_certificate():
    """the multipliers returned in dual_out (reference row scaling) certify stationarity of the restated model"""
    cfg = W.PlannerConfig()
    batch = W.make_forest_batch(64, K=40, cfg=cfg)
    agents = [3, 20]
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    st = np.ascontiguousarray(batch.state[agents]); goal = np.ascontiguousarray(batch.goal[agents]); lim = np.ascontiguousarray(batch.limits[agents])
    ctrl, cost, status, iters, kkt, dual = emul.solve_batch(batch.cfg, 2, st, goal, lim, None, off, normals, rhs, want_dual=True)
    M, D = cfg.M, cfg.dim
    for i, a in enumerate(agents):
        qp = oracle_qp_from_planes(batch, a, normals[off[i]:off[i + 1]], rhs[off[i]:off[i + 1]])
        x = ctrl[i]
        grad = 2 * qp.P @ x + qp.q
        K = off[i + 1] - off[i]
        lsc = dual[i, :40 * M * 6].reshape(40, M, 6)
        box = dual[i, 40 * M * 6:].reshape(D * M * 6, 6)
        # rebuild sum_r lam_r a_r in the reference's row order / scaling
        r = 0
        acc = np.zeros_like(grad)
        for oi in range(K):
            for m in range(M):
                for j in range(6):
                    if m == 0 and j < 3:
                        continue
                    acc += lsc[oi, m, j] * qp.G[r]          # rows n.x >= b
                    r += 1
        for k in range(D):
            for m in range(M):
                for j in range(5):
                    if m == 0 and j < 2:
                        continue
                    v = k * M * 6 + m * 6 + j
                    acc -= box[v, 2] * qp.G[r]; acc -= box[v, 3] * qp.G[r + 1]     # rows a.x <= vmax
                    r += 2
                for j in range(4):
                    if m == 0 and j < 1:
                        continue
                    v = k * M * 6 + m * 6 + j
                    acc -= box[v, 4] * qp.G[r]; acc -= box[v, 5] * qp.G[r + 1]
                    r += 2
        assert r == qp.G.shape[0]
        acc += box[:, 0] - box[:, 1]                         # variable bounds lb <= x <= ub
        # project on the null space of the equalities (their multipliers are free)
        _, S, Vt = np.linalg.svd(qp.Aeq, full_matrices=True)
        Z = Vt[int((S > 1e-10 * S[0]).sum()):].T
        res = Z.T @ (grad - acc)
        assert np.abs(res).max() < 1e-6 * max(1.0, np.abs(grad).max()), np.abs(res).max()
        assert dual[i].min() >= 0
        assert abs(cost[i] - (x @ qp.P @ x + qp.q @ x + qp.c0)) < 1e-8 * max(1.0, abs(cost[i]))


@pytest.mark.parametrize("generator,M,dim", [(0, 5, 3), (1, 10, 2), (1, 5, 3), (2, 5, 3)])
def test_assembly_kernel_matches_oracle(generator, M, dim):
    cfg = W.PlannerConfig(M=M, dim=dim)
    b = W.make_forest_batch(32, K=12, cfg=cfg)
    if generator == 1:
        near_goals(b)
    n = 6
    off = np.ascontiguousarray(b.obs_offsets[:n + 1])
    sk = off[n]
    normals, rhs = emul.assemble(b.cfg, generator, n, b.own_traj[:n].copy(), b.agent_meta[:n].copy(), b.goal[:n].copy(), off,
                                 b.obs_traj()[:sk].copy(), b.obs_meta()[:sk].copy(), b.obs_goal()[:sk].copy(),
                                 b.obs_position()[:sk].copy())
    _, n_ref, r_ref = oracle_planes(b, list(range(n)), generator)
    assert np.abs(normals - n_ref).max() <= 2.4e-7
    assert np.abs(rhs - r_ref).max() <= 2e-6
    assert (normals == n_ref).mean() > 0.99


def test_reciprocal_rsfc_generator_and_slack_mode():
    """a8': generateReciprocalRSFC (traj_planner.cpp:581-609, with obstacleSizePredictionWithConstAcc :321-358) on the
    emulator against the oracle, and the QP of mode reciprocal_rsfc: its SlackMode::COLLISIONCONSTRAINT gives every LSC row
    a free slack without a cost term (traj_optimizer.cpp:272-283, 423-425), so the oracle's model WITH the slack columns has
    the same control points as the kernel, which drops the rows; z bounds of segment 0 are +-100 (:255-258)"""
    cfg = W.PlannerConfig(M=5, dim=3, planner_mode=4)
    b = W.make_forest_batch(32, K=6, cfg=cfg)
    cfgo = oracle_config(cfg)
    n = 4
    off = np.ascontiguousarray(b.obs_offsets[:n + 1]); sk = off[n]
    # predicted sizes: every obstacle with max_acc 2, uncertainty horizon 1 s, velocity guard of the agent
    sizes = np.zeros((sk, cfg.M, 6))
    for a in range(n):
        vg = 1.0 * float((b.state[a, 3:6].astype(np.float64) ** 2).sum()) / b.limits[a, 3]
        for j in range(off[a], off[a + 1]):
            sizes[j] = orc.obstacle_sizes(cfgo, float(b.obs_meta()[j, 0]), 2.0, 1.0, vg)
    assert sizes[0, 0, 0] > 0.15 and sizes[0, 4, 5] > sizes[0, 0, 0] + 0.5        # grows by 1/2 a t^2 over the horizon
    normals, rhs = emul.assemble(cfg, 3, n, b.own_traj[:n].copy(), b.agent_meta[:n].copy(), b.goal[:n].copy(), off,
                                 b.obs_traj()[:sk].copy(), b.obs_meta()[:sk].copy(), b.obs_goal()[:sk].copy(),
                                 b.obs_position()[:sk].copy(), obs_size=sizes)
    from common import oracle_agent
    for a in range(n):
        sl = slice(off[a], off[a + 1])
        pt, nr, d = orc.generate_lsc(cfgo, orc.GEN_RSFC, oracle_agent(b, a), b.own_traj[a], b.obs_traj()[sl], b.obs_meta()[sl, 0],
                                     b.obs_meta()[sl, 1], b.obs_goal()[sl], b.obs_position()[sl], obs_size=sizes[sl])
        n_ref, r_ref = orc.pack_planes(cfgo, pt, nr, d)
        assert np.abs(normals[sl] - n_ref).max() <= 2.4e-7 and np.abs(rhs[sl] - r_ref).max() <= 2e-6
        # the QP: oracle model with K * M slack columns vs the kernel
        qp = orc.qp_build(cfgo, oracle_agent(b, a), pt, nr, d)
        assert qp.q.size == 90 + (off[a + 1] - off[a]) * cfg.M
        xe, ok = oracle_solution(qp)
        ctrl, cost, status, iters, kkt, _ = emul.solve_batch(cfg, 1, b.state[a:a + 1].copy(), b.goal[a:a + 1].copy(), b.limits[a:a + 1].copy(), None,
                                                             np.array([0, off[a + 1] - off[a]], np.int32), np.ascontiguousarray(normals[sl]),
                                                             np.ascontiguousarray(rhs[sl]))
        assert status[0] == 0 and np.abs(ctrl[0] - xe[:90]).max() < (1e-5 if ok else 1e-4), (a, np.abs(ctrl[0] - xe[:90]).max())


def test_step_kernel_matches_oracle():
    for M, dim in ((5, 3), (10, 2)):
        cfg = W.PlannerConfig(M=M, dim=dim)
        batch = W.make_forest_batch(64, K=4, cfg=cfg)
        cfgo = oracle_config(cfg)
        n = 19                                                    # not a multiple of the warps per CTA
        traj = batch.own_traj[:n].astype(np.float64) + np.random.default_rng(2).normal(0, 1e-3, batch.own_traj[:n].shape)
        ctrl = np.ascontiguousarray(np.transpose(traj, (0, 3, 1, 2))[:, :dim].reshape(n, -1))
        # failsafe (traj_planner.cpp:767-797): agents 3 and 7 did not solve and keep initial_traj
        status = np.zeros(n, np.int32); status[[3, 7]] = [2, 1]
        fallback = np.ascontiguousarray(batch.own_traj[:n])
        for st_, fb_ in ((None, None), (status, fallback)):
            got_traj, got_state, got_shift = emul.step(batch.cfg, n, ctrl, 0.1, st_, fb_)
            for a in range(n):
                want = traj[a].astype(np.float32)
                if dim == 2:
                    want[..., 2] = np.float32(cfg.z_2d)
                if st_ is not None and status[a] != 0:
                    want = fallback[a]
                assert np.array_equal(got_traj[a], want)
                st = orc.get_state_at(cfgo, want, 0.1)
                if dim == 2:
                    st[2] = np.float32(cfg.z_2d)
                assert np.allclose(got_state[a], st, rtol=2e-6, atol=2e-6)
                assert np.array_equal(got_shift[a], orc.shift_traj(cfgo, want))


@pytest.mark.parametrize("world,n_total,M,dim", [(2, 37, 5, 3), (3, 20, 10, 2), (1, 9, 5, 3), (8, 64, 5, 3)])
def test_peer_exchange_publishes_every_shard_to_every_rank(world, n_total, M, dim):
    """the fused step + all-gather of the sharded closed loop (lscqp_step_exchange / lscqp_exchange_begin), `world` ranks
    simulated on the emulator: after a step every rank holds the shifted trajectory and the new state of every agent,
    identical to the unsharded step kernel's; the double-buffered inbox survives consecutive steps; failsafe uses are
    counted; a rank that never publishes makes the others time out (counter raised, arrays untouched) instead of hanging"""
    cfg = W.PlannerConfig(M=M, dim=dim)
    batch = W.make_forest_batch(max(n_total, 16), K=4, cfg=cfg)
    sim = emul.ExchangeSim(cfg, world, n_total)
    rng = np.random.default_rng(world)
    fallback = np.ascontiguousarray(batch.own_traj[:n_total])
    for it in range(3):
        traj = batch.own_traj[:n_total].astype(np.float64) + rng.normal(0, 1e-3, batch.own_traj[:n_total].shape)
        ctrl = np.ascontiguousarray(np.transpose(traj, (0, 3, 1, 2))[:, :dim].reshape(n_total, -1))
        status = np.zeros(n_total, np.int32); status[it] = 3
        sim.step(ctrl, cfg.dt, status, fallback)
        _, want_state, want_shift = emul.step(cfg, n_total, ctrl, cfg.dt, status, fallback)
        for r in range(world):
            assert np.array_equal(sim.traj[r], want_shift), (it, r)
            assert np.array_equal(sim.state[r], want_state), (it, r)
            assert sim.counters(r)[:2] == (it + 1, 0)
    assert sum(sim.counters(r)[2] for r in range(world)) == 3
    if world > 1:
        before = sim.traj.copy()
        sim.step(ctrl, cfg.dt, status, fallback, skip_rank=world - 1)
        for r in range(world - 1):
            assert sim.counters(r)[1] == 1 and np.array_equal(sim.traj[r], before[r])


@pytest.mark.parametrize("M,dim,K,max_obs,mode", [(5, 3, 10, 40, 1), (10, 2, 9, 40, 1), (10, 2, 9, 10, 1),
                                                  (5, 3, 10, 40, 0), (10, 2, 9, 10, 2)])
def test_solve_kernel_with_communication_range_rows(M, dim, K, max_obs, mode):
    """communication-range rows (traj_optimizer.cpp:477-500; launch/simulation.launch sets range 3): dense instance
    (max_obs <= 10 selects the compact 128-thread instance of the 2-D configurations).  The rows have no mode condition
    in the reference -- its defaults are mode dlsc + range 3 (param.cpp:117,129) -- so DLSC (0) and BVC (2) run them too."""
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=mode, comm_range=0.7, max_obs=max_obs)   # tight enough to bind within the 1 s horizon
    batch = W.make_forest_batch(64, K=K, cfg=cfg)
    rng = np.random.default_rng(11)
    d = rng.normal(size=(64, 3)); d[:, 2] = 0; d /= np.linalg.norm(d, axis=1, keepdims=True)
    batch.goal = (batch.state[:, :3] + 5.0 * d).astype(np.float32)            # far goals: the agent wants to leave the range
    batch.next_waypoint = (batch.state[:, :3] + rng.uniform(-0.05, 0.05, (64, 3))).astype(np.float32)
    if dim == 2:
        batch.next_waypoint[:, 2] = cfg.z_2d
    agents = [0, 9, 30]
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    st = np.ascontiguousarray(batch.state[agents]); goal = np.ascontiguousarray(batch.goal[agents]); lim = np.ascontiguousarray(batch.limits[agents])
    wp = np.ascontiguousarray(batch.next_waypoint[agents])
    ctrl, cost, status, iters, kkt, dual = emul.solve_batch(batch.cfg, len(agents), st, goal, lim, None, off, normals, rhs,
                                                            want_dual=True, next_waypoint=wp)
    assert (status == 0).all(), status
    checked = 0
    for i, a in enumerate(agents):
        qp = oracle_qp_from_planes(batch, a, normals[off[i]:off[i + 1]], rhs[off[i]:off[i + 1]])
        assert qp.G.shape[0] == (off[i + 1] - off[i]) * (6 * M - 3) + 2 * dim * (5 * M - 2) + 2 * dim * (4 * M - 1) + dim * M * (M + 3)
        cert = orc.kkt_certificate(qp, ctrl[i])
        assert cert["primal_eq"] < 1e-8 and cert["primal_ineq"] < 1e-8, cert
        xe, ok = oracle_solution(qp)
        if ok:
            checked += 1
            assert np.abs(ctrl[i] - xe).max() < 1e-5, (a, np.abs(ctrl[i] - xe).max())
    assert checked >= 2
    # the comm rows must matter for at least one of these agents (otherwise the test proves nothing)
    assert np.abs(dual[:, -2 * dim * (M * (M - 1) // 2 + M):]).max() > 1e-6


@pytest.mark.parametrize("max_obs,k_over,presolve", [(40, 45, True), (40, 45, False), (12, 13, True), (12, 13, 3)])
def test_obstacle_lists_above_capacity_are_reported_not_truncated(max_obs, k_over, presolve):
    """the reference's model takes every obstacle it is handed (traj_optimizer.cpp:400-437 loops over getObsSize()); a
    list longer than lscqp_config.max_obs is reported per agent (LSCQP_CAPACITY, ctrl = starting point) -- never cut to
    the first max_obs -- and the other agents of the batch are solved as usual"""
    cfg = W.PlannerConfig(max_obs=max_obs, presolve=presolve)
    batch = W.make_forest_batch(64, K=min(max_obs, 12), cfg=cfg)
    agents = [0, 1, 2]
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    k = off[1] - off[0]
    reps = -(-k_over // k)
    n1 = np.concatenate([normals[off[1]:off[2]]] * reps)[:k_over]; r1 = np.concatenate([rhs[off[1]:off[2]]] * reps)[:k_over]
    normals2 = np.ascontiguousarray(np.concatenate([normals[:off[1]], n1, normals[off[2]:]]))
    rhs2 = np.ascontiguousarray(np.concatenate([rhs[:off[1]], r1, rhs[off[2]:]]))
    off2 = np.array([0, k, k + k_over, 2 * k + k_over], np.int32)
    st = np.ascontiguousarray(batch.state[agents]); goal = np.ascontiguousarray(batch.goal[agents]); lim = np.ascontiguousarray(batch.limits[agents])
    warm = np.ascontiguousarray(batch.own_traj[agents])
    ref = emul.solve_batch(cfg, 3, st, goal, lim, None, off, normals, rhs, initial_traj=warm)
    got = emul.solve_batch(cfg, 3, st, goal, lim, None, off2, normals2, rhs2, initial_traj=warm)
    assert list(got[2]) == [0, 4, 0] and list(ref[2]) == [0, 0, 0]
    assert np.array_equal(got[0][0], ref[0][0]) and np.array_equal(got[0][2], ref[0][2])
    start = np.transpose(warm[1].astype(np.float64), (2, 0, 1)).reshape(-1)
    # (the first three control points of segment 0 follow the initial state, the rest is initial_traj)
    assert np.isfinite(got[0][1]).all() and np.abs(got[0][1] - start).max() < 1e-5


def _check_neighbours(state, lo, off, idx, over, K, comm_range, rows=None):
    """CSR list of agent a = exactly the reference's obstacle set (multi_sync_simulator.cpp:319-328: every other agent whose
    L-inf distance is not above the range; all others when the range is <= 0) in ascending id order; when that set is
    larger than the capacity K: its K nearest (float distances, ties at the cut allowed) and overflow = the set's size"""
    pos = state[:, :3].astype(np.float32)
    assert off[0] == 0 and len(idx) == off[-1]
    for r in (range(len(off) - 1) if rows is None else rows):
        a = lo + r
        ids = idx[off[r]:off[r + 1]]
        d = pos - pos[a]
        linf = np.abs(d).max(axis=1).astype(np.float64)
        in_range = np.ones(pos.shape[0], bool) if comm_range <= 0 else ~(linf > comm_range)
        in_range[a] = False
        want = np.where(in_range)[0]
        assert (np.diff(ids) > 0).all()
        if len(want) <= K:
            assert np.array_equal(ids, want), (a, ids, want)
            assert over[r] == 0
        else:
            assert over[r] == len(want) and len(ids) == K and in_range[ids].all()
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).astype(np.float64)
            rest = np.setdiff1d(want, ids)
            if K > 0:
                assert d2[ids].max() <= d2[rest].min() * (1 + 1e-6) + 1e-12, (a, d2[ids].max(), d2[rest].min())


@pytest.mark.parametrize("n_total,lo,n_local,K,comm", [(300, 0, 300, 40, 0.0), (257, 100, 57, 9, 3.0), (64, 0, 64, 40, 1.0),
                                                       (41, 0, 41, 40, 0.0), (200, 0, 200, 40, 2.5), (50, 10, 30, 0, 3.0),
                                                       (1500, 700, 40, 12, 0.1)])
def test_neighbour_selection_kernel(n_total, lo, n_local, K, comm):
    rng = np.random.default_rng(n_total)
    state = np.zeros((n_total, 9), np.float32)
    state[:, :3] = rng.uniform(-6, 6, (n_total, 3)).astype(np.float32)
    state[5, :3] = state[6, :3]                               # coincident agents: tie at distance zero
    if comm == 0.1:                                           # (float) 0.1 > 0.1: a neighbour at exactly (float) 0.1 is out of range
        state[lo + 1, :3] = state[lo, :3]; state[lo + 1, 0] = state[lo, 0] + np.float32(0.1)
    off, idx, over = emul.select_neighbours(n_total, lo, n_local, K, comm, state)
    _check_neighbours(state, lo, off, idx, over, K, comm)
    if comm > 0 and K > 0:
        assert (over == 0).any()                              # the ragged case is exercised (lists shorter than K, none padded)


@pytest.mark.parametrize("generator,M,dim", [(0, 5, 3), (1, 5, 3), (1, 10, 2)])
def test_fused_assembly_reads_in_place_and_prunes_exactly(generator, M, dim):
    """lscqp_assemble_lsc_fused: (i) without prune the planes equal the gathered path bit for bit; (ii) with prune the
    dropped (obstacle, segment) pairs are zero normals, the kept ones are unchanged, and the QP solutions do not move"""
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=1)
    batch = W.make_forest_batch(48, K=20, cfg=cfg)
    near_goals(batch)
    batch.own_traj = np.ascontiguousarray(batch.own_traj)
    n = 12
    full = batch
    n_ref, r_ref = emul.assemble(cfg, generator, batch.n_agents, batch.own_traj, batch.agent_meta, batch.goal, batch.obs_offsets,
                                 batch.obs_traj(), batch.obs_meta(), batch.obs_goal(), batch.obs_position())
    n0, r0 = emul.assemble_fused(cfg, generator, False, full)
    assert np.array_equal(n0, n_ref) and np.array_equal(r0, r_ref)
    n1, r1 = emul.assemble_fused(cfg, generator, True, full)
    # the split dispatch (prune kernel -> global work list -> one thread per surviving pair) writes the same planes
    n2, r2 = emul.assemble_fused(cfg, generator, 2, full)
    assert np.array_equal(n1, n2) and np.array_equal(r1, r2)
    dropped = (n1 == 0).all(axis=2) & ~(n_ref == 0).all(axis=2)
    kept = ~dropped
    assert np.array_equal(n1[kept], n_ref[kept]) and np.array_equal(r1[kept], r_ref[kept]) and (r1[dropped] == 0).all()
    if dim == 3:
        assert dropped.mean() > 0.3, dropped.mean()            # most far pairs never reach the hull enumeration
        if generator == 1:
            assert not dropped[:, M - 1].any()                  # CLSC's last segment follows another rule: never pruned
    else:
        assert not dropped.any()                                # 2-D rows drop the z term: no pruning
    sk = int(batch.obs_offsets[n])
    args = (cfg, n, np.ascontiguousarray(batch.state[:n]), np.ascontiguousarray(batch.goal[:n]), np.ascontiguousarray(batch.limits[:n]),
            None, np.ascontiguousarray(batch.obs_offsets[:n + 1]))
    c_ref = emul.solve_batch(*args, np.ascontiguousarray(n_ref[:sk]), np.ascontiguousarray(r_ref[:sk]),
                             initial_traj=np.ascontiguousarray(batch.own_traj[:n]))
    c_pr = emul.solve_batch(*args, np.ascontiguousarray(n1[:sk]), np.ascontiguousarray(r1[:sk]),
                            initial_traj=np.ascontiguousarray(batch.own_traj[:n]))
    assert (c_ref[2] == 0).all() and (c_pr[2] == 0).all()
    assert np.abs(c_ref[0] - c_pr[0]).max() < 1e-7, np.abs(c_ref[0] - c_pr[0]).max()


@pytest.mark.parametrize("M,dim,use_sfc", [(5, 3, True), (10, 2, False), (5, 3, False)])
def test_validate_kernel_matches_oracle_is_sol_valid(M, dim, use_sfc):
    """isSolValid (traj_planner.cpp:990-1045): SFC containment + dynamic limits of the state at the replanning period"""
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=0, use_sfc=use_sfc)
    batch = W.make_forest_batch(64, K=4, cfg=cfg)
    rng = np.random.default_rng(3)
    n = batch.n_agents
    traj = np.ascontiguousarray(batch.own_traj)
    cfgo = oracle_config(cfg)
    state = np.stack([orc.get_state_at(cfgo, traj[a], cfg.dt) for a in range(n)])
    state[::3, 3:6] *= rng.uniform(2.0, 6.0, (len(state[::3]), 3)).astype(np.float32)        # some beyond the limits
    state[1::5, 6:9] *= 10.0
    sfc = None
    if use_sfc:
        lo = traj.min(axis=2) - rng.uniform(-0.02, 0.3, (n, M, 3)); hi = traj.max(axis=2) + rng.uniform(-0.02, 0.3, (n, M, 3))
        sfc = np.ascontiguousarray(np.concatenate([lo, hi], axis=2).astype(np.float32))
    got = emul.validate(cfg, n, traj, np.ascontiguousarray(state), batch.limits, sfc)
    want = np.array([orc.is_sol_valid(cfgo, orc.Agent(state[a, :3], state[a, 3:6], state[a, 6:9], batch.goal[a],
                                                      max_vel=tuple(batch.limits[a, :3]), max_acc=tuple(batch.limits[a, 3:6])),
                                      traj[a], state[a], None if sfc is None else sfc[a]) for a in range(n)], np.int32)
    assert np.array_equal(got, want)
    assert 0 < want.sum() < n


def _infeasible_case():
    """an agent whose CLSC rows (far antipodal goals: crossing goal lines) admit no point: captured from a GPU sweep
    where the multipliers overflowed to NaN after 54 iterations (tests/golden/infeasible_case.npz)"""
    g = np.load(os.path.join(GOLDEN, "infeasible_case.npz"))
    cfg = W.PlannerConfig(M=int(g["M"]), dim=int(g["dim"]), planner_mode=int(g["mode"]), world_min=tuple(g["world"][:3]),
                          world_max=tuple(g["world"][3:]))
    off = np.array([0, g["normals"].shape[0]], np.int32)
    return cfg, g, off


def test_infeasible_model_is_detected_early_and_outputs_stay_finite():
    cfg, g, off = _infeasible_case()
    for warm in (True, False):
        ctrl, cost, status, iters, kkt, dual = emul.solve_batch(
            cfg, 1, g["state"][None].copy(), g["goal"][None].copy(), g["limits"][None].copy(), None, off,
            np.ascontiguousarray(g["normals"]), np.ascontiguousarray(g["rhs"]),
            initial_traj=g["own"][None].copy() if warm else None, want_dual=True)
        assert status[0] == 2 and iters[0] < 45                     # LSCQP_INFEASIBLE, long before the iteration cap
        assert np.isfinite(ctrl).all() and np.isfinite(cost).all() and np.isfinite(dual).all() and np.isfinite(kkt).all()
        assert kkt[0, 1] > 1e-6                                     # the returned point does violate rows


def test_solve_kernel_m10_d3_instance_against_oracle():
    """the largest banded instance (M = 10, 3-D: 84 reduced variables, 256-thread CTA) against the polished oracle"""
    cfg = W.PlannerConfig(M=10, dim=3, planner_mode=1)
    batch = W.make_forest_batch(48, K=12, cfg=cfg)
    near_goals(batch)
    agents = [2, 31]
    off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
    st = np.ascontiguousarray(batch.state[agents]); goal = np.ascontiguousarray(batch.goal[agents]); lim = np.ascontiguousarray(batch.limits[agents])
    ctrl, cost, status, iters, kkt, _ = emul.solve_batch(batch.cfg, len(agents), st, goal, lim, None, off, normals, rhs,
                                                         initial_traj=np.ascontiguousarray(batch.own_traj[agents]))
    assert (status == 0).all() and kkt[:, 1].max() < 1e-9
    checked = 0
    for i, a in enumerate(agents):
        xe, ok = oracle_solution(oracle_qp_from_planes(batch, a, normals[off[i]:off[i + 1]], rhs[off[i]:off[i + 1]]))
        if ok:
            checked += 1
            assert np.abs(ctrl[i] - xe).max() < 1e-5, np.abs(ctrl[i] - xe).max()
    assert checked >= 1


@pytest.mark.parametrize("M,dim,mode,gen,K", [(5, 2, 1, 0, 12), (5, 3, 2, 2, 10), (10, 2, 0, 0, 6)])
def test_solve_kernel_other_instances_against_oracle(M, dim, mode, gen, K):
    """instances the golden file does not hold: 2-D with M = 5, BVC mode with generateBVC planes, DLSC with M = 10"""
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=mode)
    batch = W.make_forest_batch(40, K=K, cfg=cfg)
    near_goals(batch)
    agents = [1, 17, 33]
    off, normals, rhs = oracle_planes(batch, agents, gen)
    st = np.ascontiguousarray(batch.state[agents]); goal = np.ascontiguousarray(batch.goal[agents]); lim = np.ascontiguousarray(batch.limits[agents])
    ctrl, cost, status, iters, kkt, _ = emul.solve_batch(batch.cfg, len(agents), st, goal, lim, None, off, normals, rhs,
                                                         initial_traj=np.ascontiguousarray(batch.own_traj[agents]))
    assert (status == 0).all() and kkt[:, 1].max() < 1e-9
    checked = 0
    for i, a in enumerate(agents):
        qp = oracle_qp_from_planes(batch, a, normals[off[i]:off[i + 1]], rhs[off[i]:off[i + 1]])
        xe, ok = oracle_solution(qp)
        if ok:
            checked += 1
            assert np.abs(ctrl[i] - xe).max() < 1e-5, (a, np.abs(ctrl[i] - xe).max())
            assert abs(cost[i] - (xe @ qp.P @ xe + qp.q @ xe + qp.c0)) < 1e-6 * max(1.0, abs(cost[i]))
    assert checked >= 2
