#!/bin/bash
# round 2, visit 8 (1 GPU): launch-shape model check at the sharded sizes of config 5 + full tests + reduce timelines
TAG=${1:-r02f}
O=gpurun_out
mkdir -p $O
python - > $O/shapes_${TAG}.txt 2>&1 <<PY
import numpy as np, motion_planning_b200 as mp
for K,T in ((65536,64),(262144,64),(262144,128),(524288,128),(1048576,128),(2097152,128)):
    m = mp.MPPI(horizon=T, samples=K, seed=0); m.goal = np.array([0.,-1.,0.])
    r = m.bench(np.zeros(3), steps=20, warmup=3, flush_l2=True, per_kernel=True)
    print(K, T, "step %.1f us rollout %.1f us" % (r["step_ms"]*1e3, r["rollout_ms"]*1e3), "state-steps/s %.3g" % (K*T/(r["step_ms"]*1e-3)), m.launch_info())
    m.close()
PY
cat $O/shapes_${TAG}.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu_${TAG}.log; tail -4 $O/pytest_gpu_${TAG}.log
timeout 120 python profiles/reduce_timeline.py mixed > $O/reduce_timeline_${TAG}.txt 2>&1
timeout 120 python profiles/reduce_timeline.py mixed 262144 128 >> $O/reduce_timeline_${TAG}.txt 2>&1; cat $O/reduce_timeline_${TAG}.txt
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-config5 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open("$O/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernels_ms")}, d["e2e"]["ms_per_step"], d["roofline"]["frac"])
PY
