#!/bin/bash
mkdir -p gpurun_out
for grp in "tests/test_kernels_gpu.py" "tests/test_forward_gpu.py"; do
  tag=$(echo "$grp" | tr ' /' '__')
  timeout 900 python -m pytest $grp -q -m gpu -p no:cacheprovider -s > gpurun_out/f_$tag.log 2>&1
  echo "== $grp : exit $?"; grep -h "rel-L2" gpurun_out/f_$tag.log | cut -c1-400 | tail -n 70; tail -n 8 gpurun_out/f_$tag.log
done
