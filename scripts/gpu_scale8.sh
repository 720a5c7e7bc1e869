#!/bin/bash
# 8-GPU visit (gpurun --gpus 8): the scaling lines N = 1, 2, 4, 8 (weak K=65536/GPU + config 5 strong + parity), one p2p parity test
TAG=${1:-r02s}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo_${TAG}.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "sharded_equals_single_gpu and mixed-p2p" 2>&1 | tail -4 > $O/pytest_multi_${TAG}.log; tail -2 $O/pytest_multi_${TAG}.log
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/scale_n${n}_${TAG}.json 2> $O/scale_n${n}_${TAG}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 50 --warmup 5 > $O/scale_n${n}_${TAG}.json 2> $O/scale_n${n}_${TAG}.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("$O/scale_n${n}_${TAG}.json").read().strip().splitlines()[-1])
    print("N=$n", {k:d.get(k) for k in ("value","ms_per_step")}, "parity", (d.get("parity") or {}).get("pass"), "e2e", d["e2e"]["ms_per_step"], d["clocks"].get("sm_mhz_per_rank"))
    c=d.get("config5") or {}
    print("   config5", {k:c.get(k) for k in ("ms_per_step","value","one_gpu_same_run_ms_per_step","shard_alone_ms_per_rank")}, "parity", (c.get("parity") or {}).get("pass"))
except Exception as ex:
    print("N=$n no line:", ex)
PY
  tail -2 $O/scale_n${n}_${TAG}.err
done
