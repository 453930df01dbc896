#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_forward_gpu.py tests/test_kernels_gpu.py tests/test_backward_kernels_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --launch-table gpurun_out/lt_infer.json > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; python -c "
import json; d=json.load(open('gpurun_out/bench_infer.json')); print('infer', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/bench_infer.err
python bench.py --mode train --steps 10 --warmup 3 --launch-table gpurun_out/lt_train.json > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train.json')); print('train', d['value'], d['ms_per_step'], d['backward_kernels'])" || tail -5 gpurun_out/bench_train.err
