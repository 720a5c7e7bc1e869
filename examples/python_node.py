#!/usr/bin/env python
"""examples/python_node.py -- the reference's MPPI node without ROS: `Controller` (control/src/mppi:296-389) on the
B200 engine, driven by a simulated diff-drive odometry source (what control/launch/mppi_pentagon.launch wires up with
rigid2d's fake encoders).  Needs a GPU.

    python examples/python_node.py [K] [T] [callbacks]
"""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import motion_planning_b200 as mp  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
waypoints = [[1, 0], [2, 1], [1, 2], [0, 2], [0, 0]]            # control/config/waypoints.yaml:1
engine = mp.MPPI(horizon=T, samples=K, seed=0)                   # the node's `self.mppi = MPPI()`, at B200 scale
node = mp.Controller(mppi=engine, waypoints=waypoints, log=print)
plant = mp.FakeDiffDrive((0.0, 0.0, 0.0), dt=engine.dt)
t0 = time.perf_counter()
for i in range(n):
    vx, wz = node.pos_cb(plant.pose)                             # one odometry message -> one MPPI step -> one Twist
    plant.step(vx, wz)
    if i % 200 == 0:
        print("callback %5d  pose (%.3f %.3f %.3f)  cmd_vel vx=%.4f wz=%.4f" % ((i,) + tuple(plant.pose) + (vx, wz)))
dt = time.perf_counter() - t0
print("%d callbacks, %.1f us per callback (K=%d, T=%d), overflow redo count %d" % (n, dt / n * 1e6, K, T, engine.stats()["refine_overflow"]))
engine.close()
