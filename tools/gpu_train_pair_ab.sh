mkdir -p gpurun_out
for PM in 0 9; do
TDET_PAIR=$PM python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$PM bench.py --gpus 2 --mode train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_n2_p$PM.json 2> gpurun_out/bench_train_n2_p$PM.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_n2_p$PM.json')); print('train2 pair=$PM', d['value'], d['ms_per_step'])" || tail -5 gpurun_out/bench_train_n2_p$PM.err
TDET_PAIR=$PM python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_n1_p$PM.json 2> gpurun_out/bench_train_n1_p$PM.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_n1_p$PM.json')); print('train1 pair=$PM', d['value'], d['ms_per_step'])" || tail -5 gpurun_out/bench_train_n1_p$PM.err
done
