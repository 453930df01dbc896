#!/bin/bash
# N GPUs: training leg (config 4) for several SM reserves (TDET_SM_RESERVE sets the grids' reserve AND NCCL_MAX_CTAS)
N=${1:-2}; shift
mkdir -p gpurun_out
for r in "$@"; do
  for rep in 1 2; do
    echo -n "reserve $r rep $rep: "
    TDET_SM_RESERVE=$r timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --mode train --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/reserve.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step  allreduce %.2f ms exposed %.2f' % (d['value'], d['ms_per_step'], d.get('allreduce_device_ms') or -1, d.get('allreduce_exposed_ms') or -1))" || tail -3 gpurun_out/reserve.err
  done
done
