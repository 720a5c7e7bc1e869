#!/bin/bash
# multi-GPU visit (gpurun --gpus N): sharded-vs-single parity tests + the scaling bench line at N
N=${1:-2}
TAG=${2:-m}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo_${TAG}.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15 > $O/pytest_multi_${TAG}.log
tail -4 $O/pytest_multi_${TAG}.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/scale_n${n}_${TAG}.json 2> $O/scale_n${n}_${TAG}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 50 --warmup 5 > $O/scale_n${n}_${TAG}.json 2> $O/scale_n${n}_${TAG}.err
  fi
  tail -c 900 $O/scale_n${n}_${TAG}.json; echo; tail -2 $O/scale_n${n}_${TAG}.err
done
