#!/bin/bash
# ONE compute-sanitizer tool per gpurun call (B200_PROFILING.md): bash tools/gpu_sanitize.sh memcheck|racecheck|synccheck [args]
TOOL=${1:-memcheck}; shift
mkdir -p gpurun_out
python tools/sanitize_case.py "$@" > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_case.py "$@" > gpurun_out/sanitize_$TOOL.log 2>&1
echo "compute-sanitizer $TOOL exit $?"; grep -c "=========" gpurun_out/sanitize_$TOOL.log; tail -8 gpurun_out/sanitize_$TOOL.log
