"""development aid: which pass solved each agent of the bench batch (dual active-set first pass), iteration histogram"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lsc_dr_planner_b200 import capi, workloads as W
from lsc_dr_planner_b200.planner import BatchPlanner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20260001
batch = W.make_forest_batch(n, K=40, seed=seed)
pl = BatchPlanner(batch.cfg, device=0)
d = pl.upload(batch)
pl.assemble_fused_device(d)
pl.solve_device(d, want_kkt=True)
torch.cuda.synchronize()
st = d.status.cpu().numpy(); it = d.iters.cpu().numpy(); kkt = d.kkt.cpu().numpy()
kl = pl.qp.last_instances(n)
print("status hist", np.bincount(st, minlength=5), "klass hist (0 = active set; else reason)", np.bincount(kl, minlength=7))
print("iters of the active-set agents: mean %.2f, percentiles 50/90/99/max" % it[kl == 0].mean(), np.percentile(it[kl == 0], [50, 90, 99, 100]))
qa = kkt[kl == 0, 2].astype(int)
print("active rows: final mean %.1f max %d; largest during the run max %d (capacity 32)" % ((qa % 64).mean(), (qa % 64).max(), (qa // 64).max()))
print("stationarity max", kkt[kl == 0, 0].max(), "primal max", kkt[kl == 0, 1].max())
for a in np.where(kl != 0)[0][:10]:
    print("  agent", a, "klass", kl[a], "status", st[a], "iters", it[a])

stream = torch.cuda.current_stream().cuda_stream
ts = []
for rep in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pl.solve_device(d, stream=stream); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("solve ms (L2 warm): min %.4f median %.4f" % (min(ts), sorted(ts)[2]))
