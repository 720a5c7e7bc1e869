#!/bin/bash
# graph vs eager launches in the device-timed loop: N=1 (twice each), then N=8 weak + config 5
TAG=${1:-r02v}
O=gpurun_out; mkdir -p $O
for mode in graph eager graph eager; do
  MPPI_B200_BENCH=$mode python - <<PY
import numpy as np, motion_planning_b200 as mp, os
m = mp.MPPI(horizon=64, samples=65536, seed=0); m.goal = np.array([0.,-1.,0.])
r = m.bench(np.zeros(3), steps=200, warmup=10, flush_l2=True, per_kernel=False)
print("N=1", os.environ["MPPI_B200_BENCH"], "step %.2f us" % (r["step_ms"]*1e3))
PY
done
for mode in graph eager; do
  MPPI_B200_BENCH=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
      bench.py --gpus 8 --steps 100 --warmup 5 > $O/scale_n8_${mode}_${TAG}.json 2> $O/scale_n8_${mode}_${TAG}.err
  python - <<PY
import json
d=json.loads(open("$O/scale_n8_${mode}_${TAG}.json").read().strip().splitlines()[-1])
c=d.get("config5") or {}
print("N=8 $mode", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "config5", c.get("ms_per_step"), "alone", max(c.get("shard_alone_ms_per_rank") or [0]))
PY
done
