#!/bin/bash
# A/B of kernel variants selected by environment variables; each line: label, img/s, ms/step.
mkdir -p gpurun_out
run() {
  echo -n "$1 : "
  env $2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --launch-table gpurun_out/lt_$1.json 2>gpurun_out/ab_$1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step  e2e %.1f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step']))" || tail -5 gpurun_out/ab_$1.err
}
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py tests/test_backward_kernels_gpu.py tests/test_train_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
run pdl1 "TDET_PDL=1"; run pdl0 "TDET_PDL=0"; run pdl1b "TDET_PDL=1"; run pdl0b "TDET_PDL=0"
python bench.py --mode train --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'])"
