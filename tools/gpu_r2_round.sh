#!/bin/bash
# round 2: a bundle of single-GPU jobs run back to back (each writes its own files under gpurun_out/)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "bottleneck_tail" 2>&1 | tail -2
bash tools/gpu_r2_ab_env.sh "TDET_FUSE_TAIL=1" "TDET_FUSE_TAIL=0"
timeout 900 python -m pytest tests/test_forward_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/t_forward.log 2>&1
echo "== forward: exit $?"; grep -h "pretrained-like\|   [CP][2-6]  cuda" gpurun_out/t_forward.log | cut -c1-200; tail -n 3 gpurun_out/t_forward.log
timeout 900 python tools/sweep_config5.py --quick --out gpurun_out/config5_quick.json > gpurun_out/config5_quick.log 2>&1
echo "== sweep quick: exit $?"; tail -5 gpurun_out/config5_quick.log | cut -c1-300
