#!/bin/bash
# round 2: what the driver runs at round end on one GPU -- the whole GPU suite, smoke, the reference arm, the bench line
mkdir -p gpurun_out
t0=$(date +%s)
timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1
echo "== pytest -m gpu: exit $? ($(( $(date +%s) - t0 )) s)"; tail -n 3 gpurun_out/final_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -2
t0=$(date +%s)
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
echo "== reference arm: exit $? ($(( $(date +%s) - t0 )) s)"; cut -c1-600 gpurun_out/final_ref.json
t0=$(date +%s)
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "== bench: exit $? ($(( $(date +%s) - t0 )) s)"; tail -2 gpurun_out/final_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/final_bench.json').read())
for k in ('value','ms_per_step','e2e','e2e_full_copy','sustained','clocks','roofline','cpu_baseline','r101_b64','train','gpu_launches'):
    print(k, '=', json.dumps(d.get(k))[:700])"
