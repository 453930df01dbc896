#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for legs in full,sustained,train sustained,train train; do
  tag=$(echo $legs | tr ',' '_')_$i
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --legs $legs > gpurun_out/legs3_$tag.json 2> gpurun_out/legs3_$tag.err
  echo -n "legs=$legs #$i: exit $? "; python -c "
import json
d=json.loads(open('gpurun_out/legs3_$tag.json').read()); t=d['train']
print('infer %.0f img/s; train %.1f img/s %.2f ms/step; host enqueue %.2f ms; clocks %s' % (d['value'], t['img_s'], t['ms_per_step'], t['host_enqueue_ms_per_step'], t.get('clocks')))" || tail -5 gpurun_out/legs3_$tag.err
done
done
