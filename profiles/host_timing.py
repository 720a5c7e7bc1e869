"""Host-side phases of mppi_step (C ABI): where the end-to-end microseconds outside the kernels go.
   python profiles/host_timing.py [K] [T] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import motion_planning_b200 as mp  # noqa: E402
from motion_planning_b200 import _capi  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
for prec in ("mixed", "f32"):
    m = mp.MPPI(horizon=T, samples=K, precision=prec, seed=0)
    lib, h = m._lib, m._h
    goal = np.array([0.0, -1.0, 0.0])
    x, u, xn = np.zeros(3), np.empty(2), np.empty(3)
    px, pu, pn = _capi.dptr(x), _capi.dptr(u), _capi.dptr(xn)
    _capi.check(lib.mppi_set_goal(h, _capi.dptr(goal)), "goal")
    for _ in range(20):
        lib.mppi_step(h, px, pu, pn)
        x[:] = xn
    out = np.zeros(4)
    lib.mppi_debug_host_timing(h, _capi.dptr(out))
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.mppi_step(h, px, pu, pn)
        x[:] = xn
    wall = (time.perf_counter() - t0) / steps * 1e6
    lib.mppi_debug_host_timing(h, _capi.dptr(out))
    m.initialize()          # the device-resident loop restarts from x0 = 0 with a zero nominal ...
    m.goal = goal           # ... towards the same goal (bench() sends the attribute, the loop above set it through the C ABI)
    dev = m.bench(np.zeros(3), steps=50, warmup=5, flush_l2=False, per_kernel=False)["step_ms"] * 1e3
    print("%s K=%d T=%d: wall %.2f us per mppi_step (warm L2, python loop); inside the call: to rollout launched %.2f, "
          "to reduce launched %.2f, to result seen %.2f, to return %.2f (sum %.2f); device-resident step %.2f us"
          % (prec, K, T, wall, out[0], out[1], out[2], out[3], out.sum(), dev))
    m.close()
