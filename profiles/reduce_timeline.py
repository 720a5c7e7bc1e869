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
ts = np.zeros((T, 8), dtype=np.uint64)
_capi.check(m._lib.mppi_debug_reduce_timestamps(m._h, ts.ctypes.data_as(C.POINTER(C.c_uint64))), "read")
ts = ts.astype(np.int64)
t0 = ts[:, 0].min()
rel = (ts - t0) / 1e3
names = ["start", "A:min+E", "B:compact", "C:resim", "D:softmin", "ticket", "finalize_end"]
print(prec, K, T, "block start spread %.2f us" % (rel[:, 0].max()))
for j in range(1, 5):
    d = rel[:, j] - rel[:, j - 1]
    print("  phase %-10s mean %.2f  max %.2f us (block %d)" % (names[j], d.mean(), d.max(), d.argmax()))
print("  all blocks done at %.2f us; last block: ticket %.2f, finalize end %.2f us" % (
    rel[:, 4].max(), rel[:, 5].max(), rel[:, 6].max()))
