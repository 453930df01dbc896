#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/train.log 2>&1
echo "== train : exit $?"; grep -vE "^\s*$" gpurun_out/train.log | tail -n 60
