#!/bin/bash
# round 2, N GPUs: which preceding leg disturbs the training leg of the combined bench line?
N=${1:-2}
mkdir -p gpurun_out
for legs in train r101,train sustained,train full,train; do
  tag=$(echo $legs | tr ',' '_')
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --legs $legs > gpurun_out/legs_$tag.json 2> gpurun_out/legs_$tag.err
  echo -n "legs=$legs: exit $? "; python -c "
import json
d=json.loads(open('gpurun_out/legs_$tag.json').read()); t=d['train']
print('infer %.0f img/s; train %.1f img/s %.2f ms/step; allreduce device %.2f ms exposed %.2f ms' % (d['value'], t['img_s'], t['ms_per_step'], t['allreduce_device_ms'] or 0, t['allreduce_exposed_ms'] or 0))" || tail -5 gpurun_out/legs_$tag.err
done
