"""Regenerates tests/golden/mission_golden.npz (BASELINE configs 1 and 3).  Run in the build container only
(it reads the reference's mission files, which do not travel to the GPU box):
    python tests/golden/make_mission_golden.py

  c1_*  config 1: missions/forest10/forest10_1.json, 10 agents, launch settings (2-D, M=10, communication range 3),
        FIRST replan through the oracle in the reference's stage order (traj_planner.cpp:117-139):
        generateCLSC -> GoalOptimizer -> TrajOptimizer (restated model + HiGHS, polished).
  c3_*  config 3: 26 independent 10-agent instances of missions/maze10_dense/maze10_k.json rolled out in closed loop by
        the oracle (LSC -> goal LP -> QP -> doStep -> shift) for R = 2..10 replans; the inputs of replan R+1 of every agent
        (first 256 of 260) and their polished oracle solutions.
Safe Flight Corridors are ON (world_use_octomap, as the shipped launches run): every mission is paired with its world
file like the reference pairs them (lexicographic order of both directories, multi_sync_simulator_node.cpp:48-49), the
corridors are built by the oracle's restatement of generateSFC (initializeSFC on the first replan, afterwards
constructSFCFromConvexHull -- the goal mode of the launches is grid_based_planner), and they enter the goal LP and the QP.
The grid MAPF layer is outside the hot path: the next waypoint is the point one metre ahead on the straight line to the
desired goal.
"""
import glob
import os
import sys

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):      # one BLAS thread per worker process
    os.environ.setdefault(_v, "1")

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from lsc_dr_planner_b200 import missions as MS  # noqa: E402
from oracle import oracle as orc  # noqa: E402

REF = "/root/reference"


def ocfg(cfg):
    return orc.Config(M=cfg.M, n=cfg.n, phi=cfg.phi, dim=cfg.dim, dt=cfg.dt, w_control=cfg.w_control,
                      w_terminal=cfg.w_terminal, planner_mode=cfg.planner_mode, use_sfc=True, comm_range=cfg.comm_range,
                      world_min=cfg.world_min, world_max=cfg.world_max, z_2d=cfg.z_2d)


def plan_agent(cfgo, mission, a, state, goal_prev, wp, own, nbr, trajs, goals_prev, positions, sfc, final=True):
    """one agent's replan through the oracle; returns (new_goal, goal_status, x or None, planes)"""
    ag = orc.Agent(state[:3], state[3:6], state[6:9], goal_prev, next_waypoint=wp, max_vel=tuple(mission.max_vel[a]),
                   max_acc=tuple(mission.max_acc[a]), radius=float(mission.radius[a]),
                   nominal_velocity=float(mission.nominal_velocity[a]), downwash=float(mission.downwash[a]))
    pt, nr, d = orc.generate_lsc(cfgo, orc.GEN_CLSC, ag, own, trajs[nbr], mission.radius[nbr], mission.downwash[nbr],
                                 goals_prev[nbr], positions[nbr])
    ar, br = orc.goal_rows(cfgo, goal_prev, wp, pt, nr, d, sfc_last=sfc[-1])
    new_goal, t, gst = orc.goal_solve(cfgo, goal_prev, wp, ar, br)
    if gst != 0:
        return goal_prev, gst, None, (pt, nr, d)
    ag.goal = new_goal
    qp = orc.qp_build(cfgo, ag, pt, nr, d, sfc)
    # the dense interior-point checker solves these models in ~20 ms; HiGHS' active-set QP solver needs 1 s to minutes on
    # the dense communication-range models and errors out on about half of them, so it is run afterwards on a subset
    # under a wall-clock limit (highs_crosscheck) and the agreement is stored in the fixture
    sol = orc.solve_dense_ipm(qp)
    if sol.status != "Optimal":
        return new_goal, 0, None, (pt, nr, d)
    x, ok = orc.polish(qp, sol, dual_tol=1e-12)
    if final:
        HIGHS_JOBS.append((qp, x))
        # third solver: the interior-point method in C (oracle/pdip_cpu.c: equalities kept, pivoted LU)
        sc = orc.solve_pdip_c(qp)
        PDIP_DIFF.append(float(np.abs(sc.x - x).max()) if sc.status == "Optimal" else np.nan)
    return new_goal, 0, (x, ok), (pt, nr, d)


HIGHS_JOBS = []          # (qp, polished x) of the recorded step, per process
PDIP_DIFF = []           # max |x_pdip_c - x| of the recorded step's QPs, per process (NaN: not solved)


def _highs_one(qp, x, q):
    sol = orc.solve_highs(qp, time_limit=25.0)
    if sol.status != "Optimal":
        q.put(np.nan); return
    xh, ok = orc.polish(qp, sol)
    q.put(float(np.abs(xh - x).max()) if ok else np.nan)


def highs_crosscheck(jobs, limit_s=30.0, width=8):
    """max |x_highs - x_ipm| per job (NaN: HiGHS failed, was not polishable, or ran over the wall-clock limit, in which
    case its process is killed)"""
    import multiprocessing as mp
    import time
    ctx = mp.get_context("fork")
    out = np.full(len(jobs), np.nan)
    for lo in range(0, len(jobs), width):
        procs = []
        for i in range(lo, min(lo + width, len(jobs))):
            q = ctx.Queue()
            p = ctx.Process(target=_highs_one, args=(jobs[i][0], jobs[i][1], q))
            p.start(); procs.append((i, p, q))
        t0 = time.time()
        for i, p, q in procs:
            p.join(max(0.1, limit_s - (time.time() - t0)))
            if p.is_alive():
                p.kill(); p.join()
            elif not q.empty():
                out[i] = q.get()
    return out


def waypoint(pos, desired, step):
    v = desired.astype(np.float64) - pos.astype(np.float64)
    dist = np.linalg.norm(v)
    return (pos + v / max(dist, 1e-9) * min(dist, step)).astype(np.float32)


def rollout(path, world_path, R, M=10, dim=2):
    """closed loop of one mission for R replans; returns the inputs of replan R+1 and its oracle solutions"""
    mission = MS.load_mission(path, dim, 1.0)
    cfg = MS.launch_config(mission, M=M, dim=dim)
    cfgo = ocfg(cfg)
    n = mission.n_agents
    boxes = MS.load_world_csv(world_path)
    world_map = orc.Map(boxes, mission.world_min, mission.world_max, 0.1, 1.0)
    sfcs = np.zeros((n, M, 6), np.float32)
    state = np.zeros((n, 9), np.float32); state[:, :3] = mission.start
    goal = mission.start.copy()                                   # agent_manager.cpp:9
    trajs = np.stack([orc.const_vel_traj(cfgo, state[a, :3], state[a, 3:6]) for a in range(n)])
    rec = None
    for step in range(R + 1):
        off, idx = MS.neighbours_linf(state[:, :3], cfg.comm_range)
        wps = np.stack([waypoint(state[a, :3], mission.goal[a], 1.0) for a in range(n)])
        new_goal = goal.copy(); new_trajs = trajs.copy(); sols = []; oks = []
        sfc_prev = sfcs.copy(); sfc_status = np.zeros(n, np.int32)
        for a in range(n):                                         # generateSFC, traj_planner.cpp:738-753 (after the LSCs)
            if step == 0:
                ok_, box = world_map.sfc_initialize(state[a, :3], float(mission.radius[a]))
                if not ok_:
                    raise RuntimeError(f"{path}: invalid initial SFC for agent {a} (the reference throws here)")
                sfcs[a, :] = box; sfc_status[a] = 1
            else:
                st_, box = world_map.sfc_from_convex_hull([trajs[a, M - 1, 5], goal[a]], wps[a], sfcs[a, M - 1].copy(), float(mission.radius[a]))
                sfcs[a, :M - 1] = sfcs[a, 1:].copy(); sfcs[a, M - 1] = box; sfc_status[a] = st_
        for a in range(n):
            nbr = idx[off[a]:off[a + 1]]
            g, gst, sol, _ = plan_agent(cfgo, mission, a, state[a], goal[a], wps[a], trajs[a], nbr, trajs, goal, state[:, :3],
                                        sfcs[a], final=(step == R))
            new_goal[a] = g
            if sol is not None:
                x, ok = sol
                t = x.reshape(dim, M, 6).transpose(1, 2, 0)
                tr = np.full((M, 6, 3), cfg.z_2d, np.float32); tr[..., :dim] = t.astype(np.float32)
                new_trajs[a] = tr
                sols.append(x); oks.append(ok)
            else:
                sols.append(np.full(dim * M * 6, np.nan)); oks.append(False)       # failsafe: keep initial_traj
        if step == R:
            rec = dict(state=state.copy(), goal_prev=goal.copy(), goal=new_goal.copy(), wp=wps, own=trajs.copy(), off=off, idx=idx,
                       x=np.stack(sols), ok=np.array(oks), limits=np.concatenate([mission.max_vel, mission.max_acc,
                       mission.radius[:, None], mission.nominal_velocity[:, None]], 1), meta=np.stack([mission.radius, mission.downwash], 1),
                       world=np.array(mission.world_min + mission.world_max), highs_jobs=list(HIGHS_JOBS),
                       sfc=sfcs.copy(), sfc_prev=sfc_prev, sfc_status=sfc_status, boxes=boxes, first=np.full(n, step == 0),
                       pdip_maxdiff=np.array(PDIP_DIFF + [np.nan] * (n - len(PDIP_DIFF))))
            HIGHS_JOBS.clear(); PDIP_DIFF.clear()
            break
        goal = new_goal
        state = np.stack([orc.get_state_at(cfgo, new_trajs[a], cfg.dt) for a in range(n)])      # AgentManager::doStep
        trajs = np.stack([orc.shift_traj(cfgo, new_trajs[a]) for a in range(n)])                 # traj_planner.cpp:287-297
    return rec


def _roll(args):
    return rollout(*args)


def main():
    out = {}
    r1 = rollout(os.path.join(REF, "missions/forest10/forest10_1.json"), os.path.join(REF, "world/forest/forest1.csv"), 0)
    jobs = r1.pop("highs_jobs")
    for k, v in r1.items():
        out["c1_" + k] = v
    files = sorted(glob.glob(os.path.join(REF, "missions/maze10_dense/*.json")))[:26]       # lexicographic, mission.cpp:18-44
    worlds = sorted(glob.glob(os.path.join(REF, "world/maze/dense/*.csv")))[:26]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        recs = pool.map(_roll, [(f, w, 2 + (i % 9)) for i, (f, w) in enumerate(zip(files, worlds))], chunksize=1)   # replan indices 3..11
    n0 = 0
    keys = ["state", "goal_prev", "goal", "wp", "own", "x", "ok", "limits", "meta", "sfc", "sfc_prev", "sfc_status", "pdip_maxdiff"]
    cat = {k: [] for k in keys}; off = [0]; idx = []; world = []; boxes = []; boxes_off = [0]
    for r in recs[:3]:
        jobs += r["highs_jobs"]
    out["highs_maxdiff"] = highs_crosscheck(jobs)          # config 1 (10 QPs) + the first three config-3 instances
    # the same models as CPLEX-LP files (what cplex.exportModel writes, traj_optimizer.cpp:45-49): config 1 and the first
    # config-3 instance, for anyone with CPLEX to close the pin; lp_index maps file k to its fixture row
    lpdir = os.path.join(HERE, "lp")
    os.makedirs(lpdir, exist_ok=True)
    for f in glob.glob(os.path.join(lpdir, "*.lp.gz")):
        os.unlink(f)
    names = orc.variable_names(10, 5, 2)
    for k, (qp, x) in enumerate(jobs[:20]):
        tag = f"c1_agent{k}" if k < 10 else f"c3_agent{k - 10}"
        orc.write_lp(qp, os.path.join(lpdir, tag + ".lp.gz"), names,
                     comment=f"lsc_dr_planner agent QP ({tag}), restated populatebyrow (src/traj_optimizer.cpp:216-514)\n"
                             f"fixture optimum objective {float(x @ qp.P @ x + qp.q @ x + qp.c0)!r}")
    for r in recs:
        for k in keys:
            cat[k].append(r[k])
        idx.append(r["idx"] + n0); off.extend((r["off"][1:] + off[-1] - r["off"][0]).tolist())
        n0 += r["state"].shape[0]; world.append(r["world"])
        boxes.append(r["boxes"]); boxes_off.append(boxes_off[-1] + r["boxes"].shape[0])
    for k in keys:
        out["c3_" + k] = np.concatenate(cat[k])
    out["c3_off"] = np.array(off, np.int32); out["c3_idx"] = np.concatenate(idx).astype(np.int32)
    out["c3_world"] = np.stack(world)
    out["c3_boxes"] = np.concatenate(boxes); out["c3_boxes_off"] = np.array(boxes_off, np.int32)   # world boxes of instance i
    np.savez_compressed(os.path.join(HERE, "mission_golden.npz"), **out)
    print("config 1: solved", int(r1["ok"].sum()), "of", len(r1["ok"]), "| config 3:", int(out["c3_ok"].sum()), "polished of", len(out["c3_ok"]),
          "failed", int(np.isnan(out["c3_x"][:, 0]).sum()), "| C PDIP agrees (<1e-6) on", int((out["c3_pdip_maxdiff"] < 1e-6).sum()), "of 260, max",
          np.nanmax(out["c3_pdip_maxdiff"]), "| HiGHS cross-check:", int(np.isfinite(out["highs_maxdiff"]).sum()), "of",
          len(out["highs_maxdiff"]), "solved, max diff", np.nanmax(out["highs_maxdiff"]))


if __name__ == "__main__":
    main()
