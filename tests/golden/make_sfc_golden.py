"""Copies two of the reference's world files (rows: centre xyz, size xyz of every box) into tests/golden/worlds.npz so that
the SFC tests run where /root/reference does not exist (the GPU box).  Run here: python tests/golden/make_sfc_golden.py"""
import os

import numpy as np

REF = "/root/reference/world"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "worlds.npz")
np.savez_compressed(out, forest1=np.loadtxt(os.path.join(REF, "forest", "forest1.csv"), delimiter=","),
                    maze1_dense=np.loadtxt(os.path.join(REF, "maze", "dense", "maze1.csv"), delimiter=","))
print("wrote", out)
