"""Worker for the multi-GPU test / demo (launched by torchrun, one rank per GPU, NCCL):
K rollouts sharded over ranks must reproduce the single-GPU result (same Philox counters)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import motion_planning_b200 as mp  # noqa: E402
from motion_planning_b200.distributed import ShardedMPPI  # noqa: E402


def main():
    K, T = int(sys.argv[1]), int(sys.argv[2])
    precision = sys.argv[3] if len(sys.argv) > 3 else "mixed"
    exchange = sys.argv[4] if len(sys.argv) > 4 else "nccl"
    overflow = len(sys.argv) > 5 and sys.argv[5] == "overflow"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    # MPPI_TEST_ONE_GPU=1: all ranks share cuda:0 (a 1-GPU lease) -- rendezvous over gloo (NCCL refuses two ranks on one device);
    # the engines, the CUDA-IPC row exchange and the split-phase API are the same code as on N GPUs, the processes' kernels are
    # time-sliced by the driver instead of running side by side
    one_gpu = os.environ.get("MPPI_TEST_ONE_GPU") == "1"
    if one_gpu:
        local = 0
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = "cpu" if one_gpu else "cuda"
    goal = np.array([0.0, -1.0, 0.0])
    # overflow mode: an absurd screening window (50 cost units) makes the candidate lists of the fp32 screen overflow on the
    # very first step and again whenever the hold-off has run out -- what happens systematically within centimetres of a goal
    # at large K.  p2p redoes such a step in fp64 inside mppi_step, nccl / host through the MPPI_ERR_RETRY round trip of
    # ShardedMPPI.get_path; the single engine of rank 0 is fp64.
    extra = dict(refine_margin=50.0) if overflow else {}
    sh = ShardedMPPI(T, K, precision=precision, seed=0, device=local, exchange=exchange, **extra)
    one = mp.MPPI(horizon=T, samples=K, precision="f64" if overflow else precision, seed=0, device=local) if rank == 0 else None
    s = np.zeros(3)
    worst = 0.0
    for it in range(24 if overflow else 4):
        s1 = sh.get_path(s, goal)
        if one is not None:
            s2 = one.get_path(s, goal)
            err = float(np.max(np.abs(sh.latest_uvec - one.latest_uvec)) / np.max(np.abs(one.latest_uvec)))
            worst = max(worst, err, float(np.max(np.abs(s1 - s2))))
        if overflow:
            # every step is compared from identical inputs (free-running loops drift apart chaotically: the soft-min amplifies
            # 1e-12 differences of the nominal by up to 1/lam per step): all ranks take rank 0's single-engine nominal
            U = torch.from_numpy(one.latest_uvec if one is not None else np.zeros((2, T))).to(dev)
            dist.broadcast(U, src=0)
            sh.mppi.latest_uvec = U.cpu().numpy()
        s = s1
    # every rank must hold the identical nominal sequence (no broadcast is ever done)
    U = torch.from_numpy(sh.latest_uvec).to(dev)
    lo, hi = U.clone(), U.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    spread = float((hi - lo).abs().max())
    if rank == 0:
        ok = worst < 1e-9 and spread == 0.0
        if overflow:
            ovf = sh.mppi.stats()["refine_overflow"]
            ok = ok and ovf >= 2                       # step 0 and the first probe after the 8-step hold-off
            print("overflow mode: %d steps, steps redone in fp64 %d" % (it + 1, ovf))
        print("DIST %s world=%d K=%d T=%d %s/%s: max rel err vs 1 GPU %.3e, rank spread %.1e" % (
            "OK" if ok else "FAIL", world, K, T, precision, exchange, worst, spread))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
