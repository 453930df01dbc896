#!/bin/bash
# round 2 ncu evidence.  ONE ncu command per call, and only after the same plain command exited 0 (one GPU):
#   L1  launch list (gpu__time_duration.sum) of one warm inference step        L2  ... of one warm training step
#   A1  --set full: fused stem+pool, layer1.0 conv1/conv2, the dual conv3      A2  --set full: fused bottleneck tail (x2)
#   C1  --set full: layer3.0 / layer3.1 conv1, conv2, conv3 (six launches)
#   B1  --set full: FPN P2 lateral, P2 / P3 3x3 output convs                   B2  --set full: first two wgrads
mkdir -p gpurun_out
PART=${1:-L1}
case "$PART" in
  L1|A1|A2|B1|C1) PLAIN="python tools/profile_step.py --steps 3" ;;
  *)           PLAIN="python tools/profile_train.py --steps 3" ;;
esac
$PLAIN > gpurun_out/plain_$PART.log 2>&1 || { tail -5 gpurun_out/plain_$PART.log; exit 1; }
FULL="--set full --clock-control none --import-source on"
case "$PART" in
  L1) ncu --nvtx --nvtx-include "tdet_step/" --metrics gpu__time_duration.sum --clock-control none --csv \
          --log-file gpurun_out/launches_infer_r2.csv $PLAIN > gpurun_out/ncu_$PART.log 2>&1 ;;
  L2) ncu --nvtx --nvtx-include "tdet_step" --metrics gpu__time_duration.sum --clock-control none --csv \
          --log-file gpurun_out/launches_train_r2.csv $PLAIN > gpurun_out/ncu_$PART.log 2>&1 ;;
  A1) ncu $FULL -k regex:conv_gemm -s 0 -c 4 -f -o gpurun_out/prof_layer1_r2 python tools/profile_step.py --steps 1 > gpurun_out/ncu_$PART.log 2>&1 ;;
  A2) ncu $FULL -k regex:bottleneck_tail -s 0 -c 2 -f -o gpurun_out/prof_tail_r2 python tools/profile_step.py --steps 1 > gpurun_out/ncu_$PART.log 2>&1 ;;
  B1) ncu $FULL -k regex:conv_gemm -s ${SKIP:-37} -c 3 -f -o gpurun_out/prof_fpn_r2 python tools/profile_step.py --steps 1 > gpurun_out/ncu_$PART.log 2>&1 ;;
  C1) ncu $FULL -k regex:conv_gemm -s ${SKIP:-11} -c 6 -f -o gpurun_out/prof_layer3_r2 python tools/profile_step.py --steps 1 > gpurun_out/ncu_$PART.log 2>&1 ;;
  B2) ncu $FULL -k regex:wgrad_gemm -s 0 -c 2 -f -o gpurun_out/prof_wgrad_r2 python tools/profile_train.py --steps 1 > gpurun_out/ncu_$PART.log 2>&1 ;;
esac
echo "ncu $PART exit $?"
tail -3 gpurun_out/ncu_$PART.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*_r2.csv 2>/dev/null
true
