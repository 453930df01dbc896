#!/bin/bash
# ncu: launch lists of one warm inference step and one warm training step (NVTX range "tdet_step"), then
# full-set captures of representative GEMM launches.  Each ncu command runs only after the plain command exited 0.
# gpurun merges at most 64 MiB back: run part A and part B in separate calls (bash tools/gpu_ncu.sh <tag> A|B);
# L = the two launch lists only.
mkdir -p gpurun_out
R=${1:-r1s3}
PART=${2:-A}
python tools/profile_step.py --steps 3 > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
python tools/profile_train.py --steps 3 > gpurun_out/plain_train.log 2>&1 || { tail -5 gpurun_out/plain_train.log; exit 1; }
if [ "$PART" = "A" ] || [ "$PART" = "L" ]; then
ncu --nvtx --nvtx-include "tdet_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer_$R.csv python tools/profile_step.py --steps 3 > gpurun_out/ncu1.log 2>&1
echo "ncu infer list exit $?"
ncu --nvtx --nvtx-include "tdet_step" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_$R.csv python tools/profile_train.py --steps 3 > gpurun_out/ncu1t.log 2>&1
echo "ncu train list exit $?"
fi
if [ "$PART" = "A" ]; then
# fused stem + max-pool, layer1.0 shortcut, layer1.0 conv1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 0 -c 3 -f -o gpurun_out/prof_layer1_$R python tools/profile_step.py --steps 1 > gpurun_out/ncu2.log 2>&1
echo "ncu layer1 exit $?"
# FPN P2 lateral (TMA-staged coarse box), P2 and P3 3x3 output convs (CTA pairs, halo patch)
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 47 -c 3 -f -o gpurun_out/prof_fpn_$R python tools/profile_step.py --steps 1 > gpurun_out/ncu3.log 2>&1
echo "ncu fpn exit $?"
fi
if [ "$PART" = "B" ]; then
ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm -s 0 -c 2 -f -o gpurun_out/prof_wgrad_$R python tools/profile_train.py --steps 1 > gpurun_out/ncu4.log 2>&1
echo "ncu wgrad exit $?"
# operand-swapped kernel: layer2 conv1 (1x1 512->128) and conv2 (3x3, halo patch)
ncu --set full --clock-control none --import-source on -k regex:conv_swap -s 4 -c 2 -f -o gpurun_out/prof_swap_$R python tools/profile_step.py --steps 1 > gpurun_out/ncu5.log 2>&1
echo "ncu swap exit $?"
fi
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
