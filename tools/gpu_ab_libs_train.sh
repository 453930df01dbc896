#!/bin/bash
# same-box A/B of two builds of the library on the training step: build/ab/lib_<name>.so are swapped in turn
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  cp build/ab/lib_$v.so torch_detection_b200/csrc/libtdet_b200.so
  echo -n "$v $rep : "
  python bench.py --mode train --steps 20 --warmup 3 --no-cpu-baseline --launch-table gpurun_out/lt_train_${v}_$rep.json 2>gpurun_out/abt_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step' % (d['value'], d['ms_per_step']), d['backward_kernels']['wgrad'])" || tail -5 gpurun_out/abt_$v.err
done
done
