#!/bin/bash
# round 2, visit 3 (1 GPU): row exchange with pointers in the argument block and concurrent row loads
TAG=${1:-r02c}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_closed_loop.py tests/test_gpu_parity.py tests/test_gpu_controller.py -m gpu -q 2>&1 | tail -30 > $O/pytest_gpu_${TAG}.log
tail -4 $O/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu --no-config5 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open("$O/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernels_ms")}, d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["refine"])
print({p:(v["ms_per_step"],v["e2e_ms"]) for p,v in d["other_precisions"].items()})
PY
tail -3 $O/bench_${TAG}.err
timeout 120 python profiles/reduce_timeline.py mixed > $O/reduce_timeline_${TAG}.txt 2>&1; tail -25 $O/reduce_timeline_${TAG}.txt
timeout 120 python profiles/reduce_timeline.py f32 >> $O/reduce_timeline_${TAG}.txt 2>&1; tail -8 $O/reduce_timeline_${TAG}.txt
