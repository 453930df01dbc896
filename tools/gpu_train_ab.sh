#!/bin/bash
# training-step A/B over environment knobs: label, img/s, ms/step, backward kernel ms
mkdir -p gpurun_out
run() {
  echo -n "$1 : "
  env $2 python bench.py --mode train --steps 20 --warmup 3 --no-cpu-baseline --launch-table gpurun_out/lt_train_$1.json 2>gpurun_out/abt_$1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step' % (d['value'], d['ms_per_step']), d['backward_kernels']['wgrad'])" || tail -5 gpurun_out/abt_$1.err
}
run st0 "TDET_WGRAD_SMALL_TILE=0"; run st1 "TDET_WGRAD_SMALL_TILE=1"; run st0b "TDET_WGRAD_SMALL_TILE=0"; run st1b "TDET_WGRAD_SMALL_TILE=1"
