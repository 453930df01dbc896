#!/bin/bash
# A/B of kernel variants selected by environment variables; each line: label, img/s, ms/step.
mkdir -p gpurun_out
run() {
  echo -n "$1 : "
  env $2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-legs --launch-table gpurun_out/lt_$1.json 2>gpurun_out/ab_$1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step  e2e %.1f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step']))" || tail -5 gpurun_out/ab_$1.err
}
run vs0a "TDET_VARIANT_SET=0"
for v in 2 4 8 16 32 64; do run vs$v "TDET_VARIANT_SET=$v"; done
run vs0b "TDET_VARIANT_SET=0"
