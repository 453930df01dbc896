#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_infer_n$N.json 2> gpurun_out/bench_infer_n$N.err; tail -c 1500 gpurun_out/bench_infer_n$N.json; tail -3 gpurun_out/bench_infer_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err; tail -c 1500 gpurun_out/bench_train_n$N.json; tail -3 gpurun_out/bench_train_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --depth 101 --batch $((64 / N)) --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r101_n$N.json 2> gpurun_out/bench_r101_n$N.err; tail -c 600 gpurun_out/bench_r101_n$N.json; tail -3 gpurun_out/bench_r101_n$N.err
