#!/bin/bash
# 8 GPUs: multi-rank timelines (weak + config 5) with the two-stage exchange, then the scaling lines N=4, 8
TAG=${1:-r02u}
O=gpurun_out; mkdir -p $O
timeout 200 python profiles/reduce_timeline.py mixed > $O/reduce_timeline_${TAG}.txt 2>&1; cat $O/reduce_timeline_${TAG}.txt
for cfg in "" "2097152 128"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 profiles/reduce_timeline_multi.py $cfg 2>&1 | grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" >> $O/timeline_multi_${TAG}.txt
done
cat $O/timeline_multi_${TAG}.txt
for n in 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 50 --warmup 5 > $O/scale_n${n}_${TAG}.json 2> $O/scale_n${n}_${TAG}.err
  python - <<PY
import json
d=json.loads(open("$O/scale_n${n}_${TAG}.json").read().strip().splitlines()[-1])
print("N=$n", {k:d.get(k) for k in ("value","ms_per_step")}, "parity", (d.get("parity") or {}).get("pass"), "e2e", d["e2e"]["ms_per_step"])
c=d.get("config5") or {}
print("   config5", {k:c.get(k) for k in ("ms_per_step","value","one_gpu_same_run_ms_per_step","shard_alone_ms_per_rank")}, "parity", (c.get("parity") or {}).get("pass"))
PY
done
