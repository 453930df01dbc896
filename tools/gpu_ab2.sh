run() {
  echo -n "$1 : "
  env $2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs --launch-table gpurun_out/lt_$1.json 2>gpurun_out/ab_$1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step' % (d['value'], d['ms_per_step']))" || tail -5 gpurun_out/ab_$1.err
}
run r4 "TDET_RES_VARIANT=1 TDET_RES1_RING=4"; run r3 "TDET_RES_VARIANT=1 TDET_RES1_RING=3"; run r4b "TDET_RES_VARIANT=1 TDET_RES1_RING=4"; run r3b "TDET_RES_VARIANT=1 TDET_RES1_RING=3"
python - <<'PY'
import json
for t in ['r4','r3']:
    d=json.load(open('gpurun_out/lt_%s.json'%t)); print(t,[round(x['ms'],3) for x in d if x['k']==64 and x['n']==256])
PY
