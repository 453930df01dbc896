#!/bin/bash
# Runs each GPU kernel-test group in its own process (a trapped kernel poisons the CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
for grp in im2col_tile "conv_op and 1x1" "conv_op and 3x3" upsample_add prep_and_stem maxpool fold_bn unsupported; do
  tag=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" -p no:cacheprovider > gpurun_out/k_$tag.log 2>&1
  echo "== $grp : exit $?"; tail -n 25 gpurun_out/k_$tag.log
done
