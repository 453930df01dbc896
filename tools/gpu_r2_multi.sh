#!/bin/bash
# round 2, N GPUs of one box: hardware multi-GPU parity tests, then the driver's bench line at N (all legs)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/t_multi_gpu_n$N.log 2>&1
echo "== multi-GPU parity tests: exit $?"; grep -h "^{" gpurun_out/t_multi_gpu_n$N.log | cut -c1-600; tail -n 3 gpurun_out/t_multi_gpu_n$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
echo "== bench N=$N exit $?"; tail -3 gpurun_out/bench_r2_n$N.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_r2_n$N.json').read())
print('value %.1f img/s  e2e %.1f  sustained %s' % (d['value'], d['e2e']['value'], d['sustained'] and round(d['sustained']['value'],1)))
print('r101_b64', d['r101_b64']); print('train', d['train'])"
