"""GoalOptimizer (src/goal_optimizer.cpp): oracle restatement against an independent LP solve (HiGHS), the goal LP
kernel on the CPU thread emulator against the oracle, and (GPU) through the C ABI."""
import numpy as np
import pytest

from common import near_goals, oracle_config, oracle_lsc, oracle_planes
from lsc_dr_planner_b200 import workloads as W
from oracle import oracle as orc


def _case(M, dim, K, use_sfc, seed):
    """a forest batch in closed-loop-like state: goal = previous current goal, waypoint somewhere behind it"""
    cfg = W.PlannerConfig(M=M, dim=dim, planner_mode=1, use_sfc=use_sfc)
    batch = W.make_forest_batch(48, K=K, cfg=cfg, seed=seed)
    rng = np.random.default_rng(seed)
    n = batch.n_agents
    last = batch.own_traj[:, -1, -1, :]
    # g = agent.current_goal_point of the previous replan -- the same point generateCLSC builds the last-segment plane
    # from (traj_planner.cpp:692-693), close to the end of the previous solution in closed loop -- so g satisfies the
    # rows whenever the two goal lines are a safe distance apart; the next waypoint lies further out and usually
    # violates some rows, which makes the optimum t* interior.  A few agents get an unrelated goal (often infeasible).
    near_goals(batch, seed=seed, spread=0.4)
    goal = batch.goal.copy()
    wp = (goal + rng.uniform(-1.5, 1.5, (n, 3))).astype(np.float32)
    goal[3::5] = (wp[3::5] + rng.uniform(-1.5, 1.5, (len(goal[3::5]), 3))).astype(np.float32)
    goal[::7] = wp[::7]                                                         # early-out cases (g == w)
    if dim == 2:
        wp[:, 2] = cfg.z_2d; goal[:, 2] = cfg.z_2d
    sfc = None
    if use_sfc:
        lo = last - rng.uniform(0.05, 0.8, (n, 3)); hi = last + rng.uniform(0.05, 0.8, (n, 3))
        sfc = np.zeros((n, M, 6), np.float32)
        sfc[:, :, :3] = lo[:, None, :]; sfc[:, :, 3:] = hi[:, None, :]
    return batch, goal, wp, sfc


def _oracle_goal(batch, a, goal, wp, sfc, generator):
    cfgo = oracle_config(batch.cfg)
    pt, nr, d = oracle_lsc(batch, a, generator)
    ar, br = orc.goal_rows(cfgo, goal[a], wp[a], pt, nr, d, None if sfc is None else sfc[a, -1])
    out, t, st = orc.goal_solve(cfgo, goal[a], wp[a], ar, br)
    return ar, br, out, t, st


@pytest.mark.parametrize("M,dim,K,use_sfc", [(5, 3, 12, False), (10, 2, 9, True), (5, 3, 40, True)])
def test_goal_oracle_matches_highs_lp(M, dim, K, use_sfc):
    """the closed form of the restated LP agrees with HiGHS on the same rows (value and feasibility)"""
    batch, goal, wp, sfc = _case(M, dim, K, use_sfc, 11)
    n_feas = n_inf = 0
    for a in range(batch.n_agents):
        ar, br, out, t, st = _oracle_goal(batch, a, goal, wp, sfc, orc.GEN_CLSC)
        if np.linalg.norm(goal[a].astype(float) - wp[a].astype(float)) < 1e-5:
            assert st == 0 and (out == wp[a]).all() and t == 0.0
            continue
        assert len(ar) == (2 * dim if use_sfc else 0) + K
        th, feas = orc.goal_solve_highs(ar, br)
        # strict feasibility classes only (HiGHS' and the closed form's tolerances differ at the boundary)
        slack = (ar * t + br).min() if len(ar) else 1.0
        if feas and slack > -1e-7:
            assert st == 0 and abs(th - t) < 1e-7, (a, th, t)
            n_feas += 1
        elif not feas and slack < -1e-5:
            assert st == 2
            n_inf += 1
    assert n_feas > 10


@pytest.mark.parametrize("M,dim,K,use_sfc,gen", [(5, 3, 12, False, orc.GEN_CLSC), (10, 2, 9, True, orc.GEN_CLSC),
                                                 (5, 3, 40, True, orc.GEN_LSC)])
def test_goal_kernel_on_emulator_matches_oracle(M, dim, K, use_sfc, gen):
    import emul
    batch, goal, wp, sfc = _case(M, dim, K, use_sfc, 23)
    agents = list(range(batch.n_agents))
    off, normals, rhs = oracle_planes(batch, agents, gen)
    out, t, status = emul.goal(batch.cfg, len(agents), goal, wp, sfc, off, normals, rhs)
    _check_against_oracle(batch, goal, wp, sfc, gen, out, t, status)


def _check_against_oracle(batch, goal, wp, sfc, gen, out, t, status):
    for a in range(batch.n_agents):
        ar, br, o_out, o_t, o_st = _oracle_goal(batch, a, goal, wp, sfc, gen)
        slack = (ar * o_t + br).min() if len(ar) else 1.0
        if abs(slack + 1e-6) > 1e-9:                       # away from the feasibility tolerance itself
            assert status[a] == o_st, (a, status[a], o_st, slack)
        assert abs(t[a] - o_t) < 1e-12 * max(1.0, abs(o_t)) + 1e-13, (a, t[a], o_t)
        # float result: identical unless (float) t rounds differently (1 ulp of the product)
        assert np.abs(out[a] - o_out).max() <= 4e-7 * max(1.0, np.abs(o_out).max()), (a, out[a], o_out)


@pytest.mark.gpu
@pytest.mark.parametrize("M,dim,K,use_sfc,gen", [(5, 3, 12, False, orc.GEN_CLSC), (10, 2, 9, True, orc.GEN_CLSC),
                                                 (5, 3, 40, True, orc.GEN_LSC)])
def test_goal_kernel_gpu_parity(M, dim, K, use_sfc, gen):
    """lscqp_goal_host and lscqp_goal_batch (on planes assembled by the device kernel) against the oracle"""
    import torch
    from lsc_dr_planner_b200 import capi
    from lsc_dr_planner_b200.planner import BatchPlanner
    batch, goal, wp, sfc = _case(M, dim, K, use_sfc, 23)
    n = batch.n_agents
    off, normals, rhs = oracle_planes(batch, list(range(n)), gen)
    qp = capi.LscQp(batch.cfg, device=0)
    out = np.zeros((n, 3), np.float32); t = np.zeros(n); status = np.zeros(n, np.int32)
    qp.goal_host(n, goal, wp, sfc, off, normals, rhs, out, status, t_out=t)
    _check_against_oracle(batch, goal, wp, sfc, gen, out, t, status)
    # device path: planes from the assembly kernel, goal LP on the same stream
    planner = BatchPlanner(batch.cfg, device=0)
    d = planner.upload(batch)
    planner.assemble_device(d, gen)
    dev = torch.device("cuda", 0)
    g_d = torch.from_numpy(goal).to(dev); w_d = torch.from_numpy(wp).to(dev)
    sfc_d = torch.from_numpy(sfc).to(dev) if sfc is not None else None
    out_d = torch.zeros((n, 3), dtype=torch.float32, device=dev); st_d = torch.zeros(n, dtype=torch.int32, device=dev)
    t_d = torch.zeros(n, dtype=torch.float64, device=dev)
    planner.qp.goal_batch(n, g_d, w_d, sfc_d, d.obs_offsets, d.normals, d.rhs, out_d, st_d, t_out=t_d)
    torch.cuda.synchronize()
    # device-assembled normals may differ from the oracle's by one float ulp: compare at that level
    o2, t2, s2 = out_d.cpu().numpy(), t_d.cpu().numpy(), st_d.cpu().numpy()
    assert np.abs(t2 - t).max() < 1e-5 and np.abs(o2 - out).max() < 1e-5
    assert (s2 != status).sum() <= 1
