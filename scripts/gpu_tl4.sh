#!/bin/bash
TAG=${1:-r02t}; N=${2:-4}
O=gpurun_out; mkdir -p $O
for cfg in "" "2097152 128"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 profiles/reduce_timeline_multi.py $cfg 2>&1 | grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" >> $O/timeline_multi_${TAG}.txt
done
cat $O/timeline_multi_${TAG}.txt
