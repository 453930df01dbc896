#!/bin/bash
# per-launch sweep: every arm writes its in-situ launch table (two repetitions); tools/analyze_sweep_lt.py picks, per
# launch, the arms that beat the default by more than the noise
mkdir -p gpurun_out/sweep_lt
i=0
for rep in 1 2; do
  i=0
  for arm in "TDET_X=0" "$@"; do
    i=$((i+1))
    env $arm python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs --launch-table gpurun_out/sweep_lt/arm${i}_$rep.json > /dev/null 2>gpurun_out/sweep_lt/err.txt || { echo "arm $arm failed"; tail -2 gpurun_out/sweep_lt/err.txt; }
    echo "$arm" > gpurun_out/sweep_lt/arm${i}.name
  done
done
ls gpurun_out/sweep_lt | wc -l
