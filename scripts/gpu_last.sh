#!/bin/bash
# last visit of the round: the whole GPU suite + smoke + contract bench line at the final commit, host timing
TAG=${1:-r02z}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/pytest_gpu_${TAG}.log; tail -3 $O/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1; tail -2 $O/smoke_${TAG}.log
timeout 600 python bench.py --steps 100 --warmup 10 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open("$O/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernels_ms")}, "warm", d["warm_l2"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["config5"]["ms_per_step"], d["config5"]["roofline_frac_rollout_kernel"])
PY
timeout 100 python profiles/host_timing.py > $O/host_timing_${TAG}.txt 2>&1; cat $O/host_timing_${TAG}.txt
