#!/bin/bash
# multi-GPU visit (gpurun --gpus N): sharded-vs-single parity tests + the scaling bench lines at 1 and N
N=${1:-2}
TAG=${2:-m}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo_${TAG}.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -25 > $O/pytest_multi_${TAG}.log
tail -6 $O/pytest_multi_${TAG}.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/scale_n${n}_${TAG}.json 2> $O/scale_n${n}_${TAG}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 50 --warmup 5 > $O/scale_n${n}_${TAG}.json 2> $O/scale_n${n}_${TAG}.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("$O/scale_n${n}_${TAG}.json").read().strip().splitlines()[-1])
    print("N=$n", {k:d.get(k) for k in ("value","ms_per_step","parity")}, "e2e", d["e2e"]["ms_per_step"])
    c=d.get("config5") or {}
    print("   config5", {k:c.get(k) for k in ("ms_per_step","value","one_gpu_same_run_ms_per_step","parity")})
except Exception as ex:
    print("N=$n no line:", ex)
PY
  tail -3 $O/scale_n${n}_${TAG}.err
done
