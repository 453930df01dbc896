#!/bin/bash
# round 2: fused bottleneck tail -- kernel tests, forward parity, same-box A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -s -k "bottleneck_tail" > gpurun_out/k_fb.log 2>&1
echo "== fused tail kernel tests: exit $?"; grep -h "^bottleneck tail" gpurun_out/k_fb.log | tail -30; grep -n "^E " gpurun_out/k_fb.log | head -20; tail -n 4 gpurun_out/k_fb.log
timeout 900 python -m pytest tests/test_forward_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/f_fb.log 2>&1
echo "== forward: exit $?"; grep -h "rel-L2 full" gpurun_out/f_fb.log | cut -c1-420; grep -n "^E " gpurun_out/f_fb.log | head; tail -n 4 gpurun_out/f_fb.log
bash tools/gpu_r2_ab_env.sh "TDET_FUSE_TAIL=1" "TDET_FUSE_TAIL=0"
