import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as g
g.build()
from common import *
from lsc_dr_planner_b200 import capi
from lsc_dr_planner_b200.planner import BatchPlanner
import emul
cfg = W.PlannerConfig()
batch = W.make_forest_batch(64, K=40, cfg=cfg)
agents = list(range(0, 16))
off, normals, rhs = oracle_planes(batch, agents, orc.GEN_LSC)
planner = BatchPlanner(batch.cfg, 0)
n = len(agents)
state = np.ascontiguousarray(batch.state[agents]); goal = np.ascontiguousarray(batch.goal[agents]); limits = np.ascontiguousarray(batch.limits[agents])
for rep in range(2):
    ctrl = np.zeros((n, 90)); cost = np.zeros(n); status = np.zeros(n, np.int32); iters = np.zeros(n, np.int32); kkt = np.zeros((n, 4))
    planner.qp.solve_host(n, state, goal, limits, None, off, normals, rhs, ctrl, cost, status, iters, kkt, None)
    if rep == 0: ctrl0 = ctrl.copy()
print("repeatable:", np.array_equal(ctrl0, ctrl))
ectrl, ecost, estatus, eiters, ekkt, _ = emul.solve_batch(batch.cfg, n, state, goal, limits, None, off, normals, rhs)
for i, a in enumerate(agents):
    sl = slice(off[i], off[i + 1])
    qp = oracle_qp_from_planes(batch, a, normals[sl], rhs[sl])
    xe, ok = oracle_solution(qp)
    print(a, "gpu st %d it %d kkt %s err %.2e | emu st %d it %d kkt %s err %.2e | gpu-emu %.2e ok %s" % (
        status[i], iters[i], np.array2string(kkt[i], precision=1), np.abs(ctrl[i] - xe).max(),
        estatus[i], eiters[i], np.array2string(ekkt[i], precision=1), np.abs(ectrl[i] - xe).max(), np.abs(ctrl[i] - ectrl[i]).max(), ok))
