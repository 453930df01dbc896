#!/bin/bash
# same-box A/B of environment switches: bash tools/gpu_r2_ab_env.sh "NAME=VAL ..." "NAME=VAL ..." ...  (each arg = one arm)
mkdir -p gpurun_out
i=0
for rep in 1 2; do
  i=0
  for arm in "$@"; do
    i=$((i+1))
    echo -n "arm $i [$arm] rep $rep: "
    env $arm python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-legs --launch-table gpurun_out/lt_arm${i}_$rep.json 2>gpurun_out/ab_arm$i.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -5 gpurun_out/ab_arm$i.err
  done
done
