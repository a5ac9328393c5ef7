#!/bin/bash
# ncu --set full capture of ONE step-kernel launch: tools/ncu_capture.sh <name> <launch-skip> <mode> [extra args of profile_step.py]
# -> gpurun_out/<name>.ncu-rep (run under gpurun, one GPU; profile_step.py must have exited 0 without ncu first)
set -e
name=$1; skip=$2; mode=${3:-fixed_hip}; shift 3 || true
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip "$skip" --launch-count 1 \
    -f -o "gpurun_out/$name" python tools/profile_step.py $((skip + 2)) "$mode" "$@" > "gpurun_out/$name.log" 2>&1
tail -2 "gpurun_out/$name.log"
