#!/bin/bash
# ncu: launch list of one step + full-set captures of representative GEMM launches.
mkdir -p gpurun_out
python tools/profile_step.py --steps 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 64 -c 64 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 2 > gpurun_out/ncu1.log 2>&1
echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 0 -c 5 -f -o gpurun_out/prof_layer1 python tools/profile_step.py --steps 1 > gpurun_out/ncu2.log 2>&1
echo "ncu layer1 exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 56 -c 2 -f -o gpurun_out/prof_fpn python tools/profile_step.py --steps 1 > gpurun_out/ncu3.log 2>&1
echo "ncu fpn exit $?"
ls -la gpurun_out/*.ncu-rep
