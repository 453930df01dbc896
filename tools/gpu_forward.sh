#!/bin/bash
mkdir -p gpurun_out
for grp in "tests/test_kernels_gpu.py -k prep_and_stem" "tests/test_forward_gpu.py"; do
  tag=$(echo "$grp" | tr ' /' '__')
  timeout 900 python -m pytest $grp -q -m gpu -p no:cacheprovider -s > gpurun_out/f_$tag.log 2>&1
  echo "== $grp : exit $?"; tail -n 60 gpurun_out/f_$tag.log
done
