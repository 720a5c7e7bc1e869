"""Phase timeline of the reduce(+finalize) kernel from %globaltimer stamps (profiling aid).
   python profiles/reduce_timeline.py [precision] [K] [T]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import motion_planning_b200 as mp  # noqa: E402
from motion_planning_b200 import _capi  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
T = int(sys.argv[3]) if len(sys.argv) > 3 else 64
m = mp.MPPI(horizon=T, samples=K, precision=prec, seed=0)
goal = np.array([0.0, -1.0, 0.0])
_capi.check(m._lib.mppi_debug_reduce_timestamps(m._h, None), "arm")
s = np.zeros(3)
for _ in range(5):
    s = m.get_path(s, goal)
ts = np.zeros((T + 1, 8), dtype=np.uint64)          # row T: the finalizer block
_capi.check(m._lib.mppi_debug_reduce_timestamps(m._h, ts.ctypes.data_as(C.POINTER(C.c_uint64))), "read")
ts = ts.astype(np.int64)
rows, fin = ts[:T], ts[T]
t0 = rows[:, 0].min()
rel = (rows - t0) / 1e3
names = ["start", "A:min+E", "B:compact", "C:resim", "D:softmin+push"]
print(prec, K, T, "row blocks pass the PDL wait within %.2f us of each other" % (rel[:, 0].max()))
for j in range(1, 5):
    d = rel[:, j] - rel[:, j - 1]
    print("  phase %-14s mean %.2f  max %.2f us (block %d)" % (names[j], d.mean(), d.max(), d.argmax()))
f = (fin - t0) / 1e3
print("  last row pushed at %.2f us" % rel[:, 4].max())
print("  finalizer block: resident at %.2f, all rows seen %.2f, filter coefficients %.2f, result published %.2f, end %.2f us" % (
    f[0], f[1], f[3], f[4], f[2]))
