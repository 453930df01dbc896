#!/bin/bash
# A/B of kernel variants selected by environment variables; each line: label, img/s, ms/step.
mkdir -p gpurun_out
run() {
  echo -n "$1 : "
  env $2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --launch-table gpurun_out/lt_$1.json 2>gpurun_out/ab_$1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step  e2e %.1f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step']))" || tail -5 gpurun_out/ab_$1.err
}
run base "TDET_RES_VARIANT=0"; run res1 "TDET_RES_VARIANT=1"; run res2 "TDET_RES_VARIANT=2"; run res3 "TDET_RES_VARIANT=3"; run ring6 "TDET_RES1_RING=6"; run res2ring6 "TDET_RES_VARIANT=2 TDET_RES1_RING=6"
