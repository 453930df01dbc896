#!/bin/bash
N=2
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --legs $LEGS > gpurun_out/legs2_$tag.json 2> gpurun_out/legs2_$tag.err
  echo -n "$tag: exit $? "; python -c "
import json
d=json.loads(open('gpurun_out/legs2_$tag.json').read()); t=d['train']
print('train %.1f img/s %.2f ms/step; host enqueue %.2f ms; allreduce exposed %.2f; r101 %.0f' % (t['img_s'], t['ms_per_step'], t['host_enqueue_ms_per_step'], t['allreduce_exposed_ms'] or 0, d['r101_b64']['img_s']))" || tail -5 gpurun_out/legs2_$tag.err
}
LEGS=r101,train; run base A=1
LEGS=train,r101; run order2 A=1
LEGS=full,sustained,r101,train; run all A=1
