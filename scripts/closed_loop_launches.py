"""development aid: per-kernel durations of two closed-loop steps (run under `ncu --profile-from-start off`)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsc_dr_planner_b200 import workloads as W
from lsc_dr_planner_b200.closed_loop import ClosedLoopSim
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)) + "/..")
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = bench.closed_loop_batch(W, n)
sim = ClosedLoopSim(b, device=0, K=40, comm_range=3.0, use_graph=False, exchange="p2p")
for _ in range(30):
    sim.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    sim.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
