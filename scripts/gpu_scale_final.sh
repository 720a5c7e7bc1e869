#!/bin/bash
# final 8-GPU visit: scaling lines N = 2, 4, 8 (N = 1 comes from the 1-GPU evidence visit) + multi-rank timeline
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/topo_${TAG}.txt 2>&1
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
      bench.py --gpus $n --steps 100 --warmup 5 > $O/scale_n${n}_${TAG}.json 2> $O/scale_n${n}_${TAG}.err
  python - <<PY
import json
d=json.loads(open("$O/scale_n${n}_${TAG}.json").read().strip().splitlines()[-1])
c=d.get("config5") or {}
print("N=$n", d["ms_per_step"], "parity", (d.get("parity") or {}).get("pass"), "e2e", d["e2e"]["ms_per_step"], "config5", c.get("ms_per_step"), "alone", max(c.get("shard_alone_ms_per_rank") or [0]), (c.get("parity") or {}).get("pass"))
PY
done
for cfg in "" "2097152 128"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 profiles/reduce_timeline_multi.py $cfg 2>&1 | grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" >> $O/timeline_multi_${TAG}.txt
done
cat $O/timeline_multi_${TAG}.txt
