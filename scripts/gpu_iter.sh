#!/bin/bash
# short GPU visit for one kernel iteration: GPU tests, bench line (no CPU leg), launch list, optional extra command
# usage (here): gpurun --timeout 600 -- 'bash scripts/gpu_iter.sh <tag> [extra command ...]'
TAG=${1:-it}; shift
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $O/pytest_gpu_${TAG}.log
tail -4 $O/pytest_gpu_${TAG}.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
python - <<PY
import json
try:
    d = json.load(open("$O/bench_${TAG}.json"))
    print("step %.2f us  e2e %.2f us  kernels %s  frac %.3f  dev %.2e cand %d" % (d["ms_per_step"] * 1e3, d["e2e"]["ms_per_step"] * 1e3,
          {k: round(v * 1e3, 2) for k, v in d["kernels_ms"].items()}, d["roofline"]["frac"], d["refine"]["max_abs_dev_fp32_vs_fp64"], d["refine"]["candidates_last_step"]))
    print({p: round(v["ms_per_step"] * 1e3, 2) for p, v in d["other_precisions"].items()})
except Exception as ex:
    print("bench failed", ex)
PY
tail -3 $O/bench_${TAG}.err
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 24 --csv \
  --log-file $O/launches_mixed_${TAG}.csv python profiles/profile_step.py mixed 65536 64 5 > $O/ncu_mixed_${TAG}.log 2>&1
grep rollout $O/launches_mixed_${TAG}.csv | tail -2 | cut -d, -f5,13-
if [ $# -gt 0 ]; then eval "$@"; fi
