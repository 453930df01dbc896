#!/bin/bash
# round 2, N GPUs: training leg (config 4) under different all-reduce settings
N=${1:-2}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --mode train --steps 20 --warmup 5 > gpurun_out/train_$tag.json 2> gpurun_out/train_$tag.err
  echo -n "$tag: exit $? "; python -c "
import json
d=json.loads(open('gpurun_out/train_$tag.json').read())
print('%.1f img/s %.2f ms/step; allreduce device %.2f ms exposed %.2f ms' % (d['img_s'], d['ms_per_step'], d['allreduce_device_ms'] or 0, d['allreduce_exposed_ms'] or 0))" || tail -5 gpurun_out/train_$tag.err
}
run avg TDET_NCCL_AVG=1 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,ENV
grep -h "MAX_CTAS\|channels\|Channel" gpurun_out/train_avg.err | head -8
run sum TDET_NCCL_AVG=0
run sum_cta4 TDET_NCCL_AVG=0 NCCL_MAX_CTAS=4 TDET_SM_RESERVE=4
run avg_again TDET_NCCL_AVG=1
