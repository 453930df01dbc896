#!/bin/bash
mkdir -p gpurun_out
for legs in train r101,train; do
  tag=$(echo $legs | tr ',' '_')
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --legs $legs > gpurun_out/legs1_$tag.json 2> gpurun_out/legs1_$tag.err
  echo -n "legs=$legs: exit $? "; python -c "
import json
d=json.loads(open('gpurun_out/legs1_$tag.json').read()); t=d['train']
print('infer %.0f img/s; train %.1f img/s %.2f ms/step; host enqueue %.2f ms; r101 %s' % (d['value'], t['img_s'], t['ms_per_step'], t['host_enqueue_ms_per_step'], d['r101_b64'] and round(d['r101_b64']['img_s'])))" || tail -5 gpurun_out/legs1_$tag.err
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
