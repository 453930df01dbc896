#!/bin/bash
# round 2: dual-source (conv3 + projection shortcut) kernel tests, forward parity, same-box A/B of the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "dual" -p no:cacheprovider -s > gpurun_out/k_dual.log 2>&1
echo "== dual kernel tests: exit $?"; grep -h "^dual" gpurun_out/k_dual.log | tail -20; tail -n 6 gpurun_out/k_dual.log
timeout 900 python -m pytest tests/test_forward_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/f_forward.log 2>&1
echo "== forward: exit $?"; grep -h "rel-L2" gpurun_out/f_forward.log | cut -c1-300 | tail -n 12; tail -n 6 gpurun_out/f_forward.log
for rep in 1 2; do
for v in 1 0; do
  echo -n "fuse_shortcut=$v rep $rep: "
  TDET_FUSE_SHORTCUT=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-legs --launch-table gpurun_out/lt_sc${v}_$rep.json 2>gpurun_out/ab_sc$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -5 gpurun_out/ab_sc$v.err
done
done
