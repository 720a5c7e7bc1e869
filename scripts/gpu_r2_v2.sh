#!/bin/bash
# round 2, visit 2 (1 GPU): flag-in-data row exchange + finalizer block at world 1: tests, bench line, reduce timeline
TAG=${1:-r02b}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -60 > $O/pytest_gpu_${TAG}.log
tail -5 $O/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open("$O/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernels_ms")}, d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["refine"])
print("config5", {k:d["config5"][k] for k in ("ms_per_step","rollout_ms","roofline_frac_rollout_kernel")})
print({p:(v["ms_per_step"],v["e2e_ms"]) for p,v in d["other_precisions"].items()})
PY
tail -3 $O/bench_${TAG}.err
timeout 120 python profiles/reduce_timeline.py mixed > $O/reduce_timeline_${TAG}.txt 2>&1; tail -25 $O/reduce_timeline_${TAG}.txt
