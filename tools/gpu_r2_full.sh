#!/bin/bash
# full GPU test tier, then the bench line (same box)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1
echo "gpu tests exit $?"; tail -3 gpurun_out/t_all.log
bash tools/gpu_r2_ab_env.sh "TDET_EPI_FAST=0 TDET_FUSE_TAIL2=0" "TDET_EPI_FAST=1 TDET_FUSE_TAIL2=0"
