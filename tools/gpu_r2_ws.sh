#!/bin/bash
# round 2: per-warp output stores (TDET_WARP_STORES): kernel + forward parity, same-box A/B
mkdir -p gpurun_out
for grp in "tests/test_kernels_gpu.py" "tests/test_forward_gpu.py"; do
  tag=$(echo "$grp" | tr ' /' '__')
  timeout 900 python -m pytest $grp -q -m gpu -p no:cacheprovider -x > gpurun_out/f_$tag.log 2>&1
  echo "== $grp : exit $?"; tail -n 4 gpurun_out/f_$tag.log
done
for rep in 1 2; do
for v in 1 0; do
  echo -n "warp_stores=$v rep $rep: "
  TDET_WARP_STORES=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-legs --launch-table gpurun_out/lt_ws${v}_$rep.json 2>gpurun_out/ab_ws$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -5 gpurun_out/ab_ws$v.err
done
done
