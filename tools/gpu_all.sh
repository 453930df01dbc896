#!/bin/bash
# kernel tests + forward tests (separate processes), then smoke + short bench with launch table.
mkdir -p gpurun_out
for grp in "tests/test_kernels_gpu.py" "tests/test_forward_gpu.py"; do
  tag=$(echo "$grp" | tr ' /' '__')
  timeout 900 python -m pytest $grp -q -m gpu -p no:cacheprovider -s > gpurun_out/f_$tag.log 2>&1
  echo "== $grp : exit $?"; grep -h "rel-L2\|scaled chain" gpurun_out/f_$tag.log | grep -v "^E\|assert" | cut -c1-330 | tail -n 25; tail -n 12 gpurun_out/f_$tag.log
done
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --launch-table gpurun_out/launch_table.json > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
