#!/bin/bash
# run-to-run variance of the training leg (config 4, one GPU): N fresh processes
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8; do
  timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/trainvar_$i.json 2> gpurun_out/trainvar_$i.err
  python -c "
import json
d=json.loads(open('gpurun_out/trainvar_$i.json').read())
print('run $i: %.1f img/s %.2f ms/step, host enqueue %.2f ms, per step' % (d['img_s'], d['ms_per_step'], d['host_enqueue_ms_per_step']), d['per_step_ms'])" || tail -3 gpurun_out/trainvar_$i.err
done
