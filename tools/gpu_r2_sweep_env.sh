#!/bin/bash
# one-factor sweep of the experiment switches on one box (bench.py --no-extra-legs, one run per arm; the default arm
# is repeated to show the drift).  Output: one line per arm.
mkdir -p gpurun_out
run() {
  echo -n "[$1] "
  env $1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-legs 2>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step' % (d['value'], d['ms_per_step']))" || tail -2 gpurun_out/sweep.err
}
run "TDET_X=0"
for a in "$@"; do run "$a"; done
run "TDET_X=0"
