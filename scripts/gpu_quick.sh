#!/bin/bash
# quick GPU visit: tests + bench + launch lists
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest_gpu_${TAG}.log
tail -5 $O/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("$O/bench_${TAG}.json"))
print({k:d[k] for k in ("value","ms_per_step","kernels_ms","refine","other_precisions")}, d["e2e"], d["roofline"]["frac"], d["clocks"])
PY
tail -3 $O/bench_${TAG}.err
for P in mixed f32 f64; do
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv \
    --log-file $O/launches_${P}_${TAG}.csv python profiles/profile_step.py $P 65536 64 5 > $O/ncu_${P}_${TAG}.log 2>&1
done
