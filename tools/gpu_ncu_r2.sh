#!/bin/bash
# round 2 ncu evidence (each ncu command only after the same plain command exited 0; one GPU):
#   L  launch lists (gpu__time_duration.sum) of one warm inference step and one warm training step
#   A  --set full captures: fused stem+pool, layer1.0 conv1/conv2, the dual-source conv3, both fused bottleneck tails
#   B  --set full captures: FPN P2 lateral, P2 / P3 3x3 output convs, first two wgrads
mkdir -p gpurun_out
PART=${1:-L}
python tools/profile_step.py --steps 3 > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
if [ "$PART" = "L" ]; then
python tools/profile_train.py --steps 3 > gpurun_out/plain_train.log 2>&1 || { tail -5 gpurun_out/plain_train.log; exit 1; }
ncu --nvtx --nvtx-include "tdet_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer_r2.csv python tools/profile_step.py --steps 3 > gpurun_out/ncu1.log 2>&1
echo "ncu infer list exit $?"
ncu --nvtx --nvtx-include "tdet_step" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_r2.csv python tools/profile_train.py --steps 3 > gpurun_out/ncu1t.log 2>&1
echo "ncu train list exit $?"
fi
if [ "$PART" = "A" ]; then
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 0 -c 4 -f -o gpurun_out/prof_layer1_r2 python tools/profile_step.py --steps 1 > gpurun_out/ncu2.log 2>&1
echo "ncu layer1 exit $?"
ncu --set full --clock-control none --import-source on -k regex:bottleneck_tail -s 0 -c 2 -f -o gpurun_out/prof_tail_r2 python tools/profile_step.py --steps 1 > gpurun_out/ncu3.log 2>&1
echo "ncu tail exit $?"
fi
if [ "$PART" = "B" ]; then
python tools/profile_train.py --steps 1 > gpurun_out/plain_train.log 2>&1 || { tail -5 gpurun_out/plain_train.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 37 -c 3 -f -o gpurun_out/prof_fpn_r2 python tools/profile_step.py --steps 1 > gpurun_out/ncu4.log 2>&1
echo "ncu fpn exit $?"
ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm -s 0 -c 2 -f -o gpurun_out/prof_wgrad_r2 python tools/profile_train.py --steps 1 > gpurun_out/ncu5.log 2>&1
echo "ncu wgrad exit $?"
fi
ls -la gpurun_out/*.ncu-rep 2>/dev/null; du -sh gpurun_out
