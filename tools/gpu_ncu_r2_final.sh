#!/bin/bash
# round 2, final build: --set full captures summarised ON the box (the four .ncu-rep files together exceed what gpurun
# copies back); only the markdown summaries and the top stall sites come home.  Usage: gpu_ncu_r2_final.sh [FPN_SKIP]
mkdir -p gpurun_out
FPN_SKIP=${1:-38}
for PART in A1 A2 B1 C1 B2; do
  case "$PART" in
    A1) REP=prof_layer1_r2 ;; A2) REP=prof_tail_r2 ;; B1) REP=prof_fpn_r2 ;; C1) REP=prof_layer3_r2 ;; B2) REP=prof_wgrad_r2 ;;
  esac
  if [ "$PART" = B1 ]; then SKIP=$FPN_SKIP bash tools/gpu_ncu_r2.sh B1 | head -1
  elif [ "$PART" = C1 ]; then SKIP=14 bash tools/gpu_ncu_r2.sh C1 | head -1
  else bash tools/gpu_ncu_r2.sh $PART | head -1; fi
  python tools/ncu_summary.py --rep gpurun_out/$REP.ncu-rep --out gpurun_out/sum_$REP.md --title "$REP" > /dev/null 2>&1
  for k in 0 1 2; do python tools/ncu_source_top.py gpurun_out/$REP.ncu-rep $k 2>/dev/null | head -14 >> gpurun_out/top_$REP.txt; done
  rm -f gpurun_out/$REP.ncu-rep
done
ls -la gpurun_out/sum_*.md gpurun_out/top_*.txt
