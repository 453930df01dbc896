#!/bin/bash
# Backward-kernel parity groups, each in its own process (a trapped kernel poisons the CUDA context).
mkdir -p gpurun_out
for grp in dgrad_stride1 dgrad_3x3_stride2 shortcut_parity "test_wgrad" backward_helpers; do
  tag=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_backward_kernels_gpu.py -q -m gpu -k "$grp" -p no:cacheprovider -s > gpurun_out/bk_$tag.log 2>&1
  echo "== $grp : exit $?"; grep -E "rel-L2|passed|failed|Error|error" gpurun_out/bk_$tag.log | tail -n 40
done
