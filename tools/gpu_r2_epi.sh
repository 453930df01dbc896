#!/bin/bash
# compile-time epilogue variants (TDET_EPI_FAST): kernel + forward tests, then a same-box A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py tests/test_guard_bands_gpu.py -x -q -m gpu > gpurun_out/t_epi.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/t_epi.log
bash tools/gpu_r2_ab_env.sh "TDET_EPI_FAST=0 TDET_FUSE_TAIL2=0" "TDET_EPI_FAST=1 TDET_FUSE_TAIL2=0" "TDET_EPI_FAST=1 TDET_FUSE_TAIL2=1"
