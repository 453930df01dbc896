#!/bin/bash
# round 2: the whole GPU suite, one pytest process per file (a trapped kernel poisons the CUDA context)
mkdir -p gpurun_out
for f in tests/test_kernels_gpu.py tests/test_forward_gpu.py tests/test_backward_kernels_gpu.py tests/test_train_gpu.py tests/test_split_precision_gpu.py tests/test_multi_gpu.py tests/test_guard_bands_gpu.py; do
  tag=$(basename $f .py)
  timeout 1500 python -m pytest $f -q -m gpu -p no:cacheprovider -s > gpurun_out/t_$tag.log 2>&1
  echo "== $f : exit $?"; grep -h "full-size\|under an 8-SM\|pretrained-like\|rel-L2 variant" gpurun_out/t_$tag.log | cut -c1-300; tail -n 3 gpurun_out/t_$tag.log
done
