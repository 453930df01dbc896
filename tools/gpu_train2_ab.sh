#!/bin/bash
# 2-GPU training A/B over environment knobs
mkdir -p gpurun_out
i=0
for cfg in "TDET_X=0" "TDET_PAIR=0" "TDET_PAIR=1" "TDET_SWAP=0"; do
  i=$((i+1))
  env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus 2 --mode train --steps 10 --warmup 3 --no-cpu-baseline --launch-table gpurun_out/lt_train2_$i.json > gpurun_out/bench_train2_$i.json 2> gpurun_out/bench_train2_$i.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train2_$i.json')); print('$cfg', d['value'], d['ms_per_step'], d['backward_kernels'])" || tail -5 gpurun_out/bench_train2_$i.err
done
