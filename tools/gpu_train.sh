#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "wgrad" > gpurun_out/bk_wgrad.log 2>&1
echo "== wgrad : exit $?"; tail -n 3 gpurun_out/bk_wgrad.log
timeout 1500 python -m pytest tests/test_train_gpu.py -x -q -m gpu -p no:cacheprovider -s > gpurun_out/train.log 2>&1
echo "== train : exit $?"; grep -vE "^\s*$" gpurun_out/train.log | tail -n 60
