"""Closed loop on the model all the way INTO the goal (the reference node stops stepping at thresh = 0.05 m): per-step
host time, candidate counts and candidate-list overflows of the mixed mode as the soft-min support grows.
   python profiles/closed_loop_probe.py [precision]"""
import os
import sys
import time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, motion_planning_b200 as mp
K, T = 65536, 64
prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
m = mp.MPPI(horizon=T, samples=K, precision=prec, seed=0)
print("precision", prec)
goal = np.array([0.0, -1.0, 0.0])
s = np.zeros(3)
lib, h = m._lib, m._h
t_hist, ovf_hist, cand_hist, dist = [], [], [], []
prev_ovf = 0
for i in range(900):
    t0 = time.perf_counter()
    s = m.get_path(s, goal)
    t_hist.append((time.perf_counter() - t0) * 1e6)
    st = m.stats()
    ovf_hist.append(st["refine_overflow"]); cand_hist.append(st["refine_candidates"]); dist.append(np.linalg.norm(s[:2] - goal[:2]))
t_hist = np.array(t_hist); ovf = np.array(ovf_hist); cand = np.array(cand_hist); dist = np.array(dist)
for a in range(0, 900, 100):
    sl = slice(a, a + 100)
    print("steps %3d-%3d: dist %.3f -> %.3f  median %.1f us  max %.1f us  overflow total %d  candidates median %d max %d" % (
        a, a + 99, dist[sl][0], dist[sl][-1], np.median(t_hist[sl]), t_hist[sl].max(), ovf[sl][-1], np.median(cand[sl]), cand[sl].max()))
