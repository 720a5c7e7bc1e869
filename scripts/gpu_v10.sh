#!/bin/bash
# 1 GPU: correctness of the two-stage exchange + timeline + bench
TAG=${1:-r02h}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu_${TAG}.log; tail -4 $O/pytest_gpu_${TAG}.log
timeout 120 python profiles/reduce_timeline.py mixed > $O/reduce_timeline_${TAG}.txt 2>&1; cat $O/reduce_timeline_${TAG}.txt
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-config5 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open("$O/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernels_ms")}, d["e2e"]["ms_per_step"], d["roofline"]["frac"])
PY
