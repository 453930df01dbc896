#!/bin/bash
# smoke + short bench + per-launch table + ncu launch list (device time per launch).
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --launch-table gpurun_out/launch_table.json > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python tools/profile_step.py --steps 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 76 -c 76 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 2 > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu.log
