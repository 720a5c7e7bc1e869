#!/bin/bash
# round 2, visit 4 (1 GPU): full GPU suite (user models, tracking, grid updates, row exchange), bench line, reduce timelines
TAG=${1:-r02d}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -70 > $O/pytest_gpu_${TAG}.log
tail -6 $O/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu --no-config5 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open("$O/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernels_ms")}, d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["refine"])
print({p:(v["ms_per_step"],v["e2e_ms"]) for p,v in d["other_precisions"].items()})
PY
tail -3 $O/bench_${TAG}.err
timeout 120 python profiles/reduce_timeline.py mixed > $O/reduce_timeline_${TAG}.txt 2>&1; tail -12 $O/reduce_timeline_${TAG}.txt
timeout 120 python profiles/reduce_timeline.py f32 >> $O/reduce_timeline_${TAG}.txt 2>&1; tail -8 $O/reduce_timeline_${TAG}.txt
