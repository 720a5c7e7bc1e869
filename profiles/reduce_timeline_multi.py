"""Per-rank phase timeline of the sharded step's reduce kernel (profiling aid): where the coupling of N ranks costs time.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/reduce_timeline_multi.py [K_total] [T]
%globaltimer is per GPU, so every rank reports durations relative to the moment ITS row blocks passed the PDL wait."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from motion_planning_b200 import _capi  # noqa: E402
from motion_planning_b200.distributed import ShardedMPPI  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536 * world
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
m = ShardedMPPI(T, K, precision="mixed", seed=0, device=local, exchange="p2p")
lib, h = m.mppi._lib, m.mppi._h
_capi.check(lib.mppi_debug_reduce_timestamps(h, None), "arm")
m.mppi.goal = np.array([0.0, -1.0, 0.0])
res = []
for rep in range(3):
    dist.barrier()
    torch.cuda.synchronize()
    r = m.mppi.bench(np.zeros(3), steps=12, warmup=3, flush_l2=True, per_kernel=False)     # the stamps of the LAST step stay
    ts = np.zeros((T + 1, 8), dtype=np.uint64)
    _capi.check(lib.mppi_debug_reduce_timestamps(h, ts.ctypes.data_as(C.POINTER(C.c_uint64))), "read")
    ts = ts.astype(np.int64)
    rows, fin = ts[:T], ts[T]
    t0 = rows[:, 0].min()
    ph = [(rows[:, j] - rows[:, j - 1]) / 1e3 for j in range(1, 5)]          # A, B, C, D (D = soft-min + exchange + merge)
    res.append([r["step_ms"] * 1e3, (rows[:, 0].max() - t0) / 1e3, (rows[:, 4].max() - t0) / 1e3, (fin[1] - t0) / 1e3, (fin[3] - t0) / 1e3,
                (fin[2] - t0) / 1e3] + [f(x) for x in ph for f in (np.mean, np.max)] + [float(r["refine_candidates"])])
out = torch.tensor(res[-1], dtype=torch.float64, device="cuda")
allr = [torch.empty_like(out) for _ in range(world)]
dist.all_gather(allr, out)
if rank == 0:
    print("world %d K_total %d T %d (us; per rank): step | row blocks start spread | own last row done | merged rows seen by the finalizer | "
          "filter coefficients | finalizer end | phases A B C D as mean,max | candidates" % (world, K, T))
    for g, v in enumerate(allr):
        v = v.tolist()
        print("  rank %d: " % g + "  ".join("%6.2f" % x for x in v[:6]) + "  |  " + "  ".join("%5.2f,%5.2f" % (v[6 + 2 * i], v[7 + 2 * i]) for i in range(4))
              + "  | %d" % v[14])
dist.barrier()
dist.destroy_process_group()
