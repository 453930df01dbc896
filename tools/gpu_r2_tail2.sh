#!/bin/bash
# planes = 128 tail kernel: kernel tests, cycle trace, then (full) forward tests and a same-box A/B of the switch
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "bottleneck_tail" -s > gpurun_out/t_tail2.log 2>&1
echo "tail tests exit $?"; tail -2 gpurun_out/t_tail2.log
timeout 120 python tools/trace_bottleneck_tail.py --planes 128 --tiles 1 --all > gpurun_out/trace_t2_v3.txt 2>&1; head -4 gpurun_out/trace_t2_v3.txt; tail -3 gpurun_out/trace_t2_v3.txt
if [ "$1" == "full" ]; then
  timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_guard_bands_gpu.py -x -q -m gpu > gpurun_out/t_fwd_tail2.log 2>&1
  echo "forward tests exit $?"; tail -3 gpurun_out/t_fwd_tail2.log
  bash tools/gpu_r2_ab_env.sh "TDET_FUSE_TAIL2=0" "TDET_FUSE_TAIL2=1"
fi
