#!/bin/bash
TAG=${1:-r02w}
O=gpurun_out; mkdir -p $O
python - <<PY
import numpy as np, motion_planning_b200 as mp
for K,T in ((65536,64),(262144,128)):
    m = mp.MPPI(horizon=T, samples=K, seed=0); m.goal = np.array([0.,-1.,0.])
    for rep in range(2):
        r = m.bench(np.zeros(3), steps=200, warmup=10, flush_l2=True, per_kernel=False)
        print(K, T, "flushed step %.2f us" % (r["step_ms"]*1e3))
    r = m.bench(np.zeros(3), steps=200, warmup=10, flush_l2=False, per_kernel=False)
    print(K, T, "warm step %.2f us" % (r["step_ms"]*1e3))
    m.close()
PY
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-config5 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open("$O/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernels_ms")}, d["e2e"]["ms_per_step"], d["e2e"]["warm_l2_ms_per_step"], d["roofline"]["frac"])
PY
