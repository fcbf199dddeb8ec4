"""Regenerates the committed golden fixtures.  Run in the build container only:
    python tests/golden/make_golden.py

  gjk_golden.npz   inputs/outputs of the REFERENCE's own openGJK (oracle/_ref/libopengjk_ref.so, compiled from
                   /root/reference/src/openGJK/openGJK.cpp by oracle/Makefile) on 6-point hulls shaped like the
                   relative control points normalVectorBetweenPolys feeds it (traj_planner.cpp:1179-1205).
  qp_golden.npz    restated-model inputs of a few agents with the HiGHS solution polished to ~1e-12
                   (CPLEX itself is absent, see oracle/lscqp_oracle.h); pins the oracle against drift.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle as orc  # noqa: E402


def gjk_cases(rng, n):
    pts = np.zeros((n, 6, 3))
    for t in range(n):
        kind = t % 5
        p = rng.normal(size=(6, 3)) * rng.uniform(0.01, 2) + rng.normal(size=3) * rng.uniform(0, 3)
        if kind == 1:      # constant-velocity relative trajectory: collinear control points
            p = rng.normal(size=3) + np.linspace(0, 1, 6)[:, None] * rng.normal(size=3)
        elif kind == 2:    # planar (2-D missions keep z equal)
            p[:, 2] = 0
        elif kind == 3:    # hover: all control points coincide
            p = np.tile(rng.normal(size=3), (6, 1))
        elif kind == 4:    # origin inside the hull (colliding hulls)
            p = rng.normal(size=(6, 3))
        pts[t] = p.astype(np.float32).astype(np.float64)   # the reference widens float control points
    return pts


def main():
    assert orc.ref_available(), "oracle/_ref/libopengjk_ref.so missing: run `make -C oracle ref` where /root/reference exists"
    rng = np.random.default_rng(20260017)
    pts = gjk_cases(rng, 600)
    v = np.zeros((pts.shape[0], 3)); d = np.zeros(pts.shape[0])
    for i in range(pts.shape[0]):
        v[i], d[i] = orc.ref_gjk(pts[i])
    np.savez_compressed(os.path.join(HERE, "gjk_golden.npz"), pts=pts, v=v, dist=d)

    from common import oracle_planes, oracle_qp_from_planes, oracle_solution
    from lsc_dr_planner_b200 import workloads as W
    out = {}
    cases = [("m5d3lsc", W.PlannerConfig(M=5, dim=3, planner_mode=1), 40), ("m10d2lsc", W.PlannerConfig(M=10, dim=2, planner_mode=1), 9),
             ("m5d3dlsc", W.PlannerConfig(M=5, dim=3, planner_mode=0), 16)]
    for name, cfg, K in cases:
        batch = W.make_forest_batch(64, K=K, cfg=cfg)
        agents = [0, 5, 11, 17, 23, 42]
        off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
        xs = []
        keep = []
        for i, a in enumerate(agents):
            qp = oracle_qp_from_planes(batch, a, normals[off[i]:off[i + 1]], rhs[off[i]:off[i + 1]])
            x, ok = oracle_solution(qp)
            if ok:
                xs.append(x); keep.append(i)
        out[name + "_agents"] = np.array([agents[i] for i in keep])
        out[name + "_off"] = off; out[name + "_normals"] = normals; out[name + "_rhs"] = rhs
        out[name + "_keep"] = np.array(keep); out[name + "_x"] = np.array(xs)
    np.savez_compressed(os.path.join(HERE, "qp_golden.npz"), **out)
    print("wrote gjk_golden.npz (600 cases) and qp_golden.npz")


if __name__ == "__main__":
    main()
