#!/bin/bash
# same-box A/B of two builds of the library: build/ab/lib_<name>.so are swapped in turn
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in "$@"; do
  cp build/ab/lib_$v.so torch_detection_b200/csrc/libtdet_b200.so
  echo -n "$v $rep : "
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-legs --launch-table gpurun_out/lt_${v}_$rep.json 2>gpurun_out/ab_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -5 gpurun_out/ab_$v.err
done
done
