#!/bin/bash
# per-launch sweep: every arm writes its in-situ launch table (two repetitions); tools/analyze_sweep_lt.py picks, per
# launch, the arms that beat the default by more than the noise
# MODE=train sweeps the training step (bench.py --mode train) into gpurun_out/sweep_lt_train
MODE=${MODE:-infer}
OUT=gpurun_out/sweep_lt; EXTRA="--no-extra-legs"
if [ "$MODE" = train ]; then OUT=gpurun_out/sweep_lt_train; EXTRA="--mode train"; fi
mkdir -p $OUT
i=0
for rep in 1 2; do
  i=0
  for arm in "TDET_X=0" "$@"; do
    i=$((i+1))
    env $arm python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA --launch-table $OUT/arm${i}_$rep.json > /dev/null 2>$OUT/err.txt || { echo "arm $arm failed"; tail -2 $OUT/err.txt; }
    echo "$arm" > $OUT/arm${i}.name
  done
done
ls $OUT | wc -l
