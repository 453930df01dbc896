#!/bin/bash
# training-step A/B over environment knobs: label, img/s, ms/step, backward kernel ms
mkdir -p gpurun_out
run() {
  echo -n "$1 : "
  env $2 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline --launch-table gpurun_out/lt_train_$1.json 2>gpurun_out/abt_$1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step' % (d['value'], d['ms_per_step']), 'dgrad %.3f' % d['backward_kernels']['dgrad']['ms'])" || tail -5 gpurun_out/abt_$1.err
}
run base "TDET_X=0"
for v in 2 4 8 16 32 64; do run vs$v "TDET_VARIANT_SET=$v"; done
for v in 1 2 3; do run rv$v "TDET_RES_VARIANT=$v"; done
run base2 "TDET_X=0"
