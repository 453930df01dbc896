#!/bin/bash
mkdir -p gpurun_out
TDET_CHUNKS="2,2" timeout 600 python -m pytest tests/test_forward_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
for c in "0" "1" "2" "4" "1,2" "2,4" "4,8" "2,4,8" "1,1" "2,2" "4,4"; do
  echo -n "chunks=$c : "
  TDET_CHUNKS="$c" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step  e2e %.1f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step']))"
done
