#!/bin/bash
# round 2: source-level ncu capture of the epilogue-bound 256-wide short-K convs (dual conv3, 3x3, conv3+residual)
mkdir -p gpurun_out
python tools/profile_step.py --steps 1 > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 3 -f -o gpurun_out/prof_epi_r2 python tools/profile_step.py --steps 1 > gpurun_out/ncu_epi.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_epi.log; ls -la gpurun_out/*.ncu-rep
