"""Tiny driver for ncu: a few MPPI steps at one precision (device-resident loop, no CPU legs).
    ncu ... python profiles/profile_step.py [precision] [K] [T] [steps]"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import motion_planning_b200 as mp  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
T = int(sys.argv[3]) if len(sys.argv) > 3 else 64
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
m = mp.MPPI(horizon=T, samples=K, precision=prec, seed=0)
m.goal = np.array([0.0, -1.0, 0.0])
r = m.bench(np.zeros(3), steps=steps, warmup=3, flush_l2=False, per_kernel=False)
print(prec, K, T, r, m.launch_info())
