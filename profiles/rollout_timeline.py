"""Timeline of one MPPI step from %globaltimer stamps of every rollout CTA and every reduce block (profiling aid).
Needs a library built with the stamps compiled in:
    python -m motion_planning_b200.build --tag=timeline -DMPPI_EXP_TIMELINE
    MPPI_B200_LIB=motion_planning_b200/lib/libmppi_b200_timeline.so python profiles/rollout_timeline.py [precision] [K] [T]
"""
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
lib, h = m._lib, m._h
li = m.launch_info()
n = (K + 63) // 64 if li["block"] == 512 else li["grid"]   # stamps per partial record (tile)
_capi.check(lib.mppi_debug_rollout_timestamps(h, None, 0), "arm rollout")
_capi.check(lib.mppi_debug_reduce_timestamps(h, None), "arm reduce")
s = np.zeros(3)
for flush in (False, True):
    for _ in range(4):
        s = m.get_path(s, goal)
    if flush:
        _capi.check(lib.mppi_debug_flush_l2(h), "flush")
    import time
    t0 = time.perf_counter()
    s = m.get_path(s, goal)
    host_us = (time.perf_counter() - t0) * 1e6
    r = np.zeros((n, 8), dtype=np.uint64)
    _capi.check(lib.mppi_debug_rollout_timestamps(h, r.ctypes.data_as(C.POINTER(C.c_uint64)), n), "read rollout")
    d = np.zeros((T, 8), dtype=np.uint64)
    _capi.check(lib.mppi_debug_reduce_timestamps(h, d.ctypes.data_as(C.POINTER(C.c_uint64))), "read reduce")
    r, d = r.astype(np.int64), d.astype(np.int64)
    if r[:, 0].max() == 0:
        sys.exit("no rollout stamps: library not built with -DMPPI_EXP_TIMELINE")
    t0 = r[:, 0].min()
    ru, du = (r[:, :7] - t0) / 1e3, (d[:, :7] - t0) / 1e3
    smid = r[:, 7]
    q = lambda a: "min %.2f  p50 %.2f  p90 %.2f  max %.2f" % (a.min(), np.percentile(a, 50), np.percentile(a, 90), a.max())  # noqa: E731
    print("%s K=%d T=%d block=%d records=%d L2 %s; host-side mppi_step %.1f us" % (prec, K, T, li["block"], n, "flushed" if flush else "warm", host_us))
    if li["block"] == 512:
        shared = np.arange(n) % 7 == 6
        print("  shared (time-cut) tiles: loop end %s" % q(ru[shared, 3]))
        print("  full tiles:              loop end %s" % q(ru[~shared, 3]))
    print("  CTA entry (after first)          ", q(ru[:, 0]))
    print("  prologue: entry -> loads issued  ", q(ru[:, 1] - ru[:, 0]))
    print("  prologue: loads -> barrier passed", q(ru[:, 2] - ru[:, 1]))
    print("  T-step loop                      ", q(ru[:, 3] - ru[:, 2]))
    print("  terminal cost + total store      ", q(ru[:, 4] - ru[:, 3]))
    print("  block barrier                    ", q(ru[:, 5] - ru[:, 4]))
    print("  transposed pass                  ", q(ru[:, 6] - ru[:, 5]))
    print("  CTA done (after first entry)     ", q(ru[:, 6]))
    per_sm = {}
    for i in range(n):
        per_sm.setdefault(int(smid[i]), []).append(i)
    cnt = np.array([len(v) for v in per_sm.values()])
    last = np.array([ru[v, 6].max() for v in per_sm.values()])
    print("  SMs used %d, CTAs per SM min %d max %d; SM finish time: %s" % (len(per_sm), cnt.min(), cnt.max(), q(last)))
    for c in sorted(set(cnt)):
        print("     SMs with %d CTAs: %d, finish %s" % (c, (cnt == c).sum(), q(last[cnt == c])))
    print("  reduce: block start %s" % q(du[:, 0]))
    print("  reduce: phases done  A %.2f  B %.2f  C %.2f  D %.2f  ticket %.2f  finalize end %.2f us" % (
        du[:, 1].max(), du[:, 2].max(), du[:, 3].max(), du[:, 4].max(), du[:, 5].max(), du[:, 6].max()))
