#!/bin/bash
# compute-sanitizer over the step path: memcheck, racecheck (the step kernel has five block
# barriers in the lane sort and two per physics iteration over shared memory that is re-used between the sort and the
# cold slots), initcheck and synccheck on tools/sanitize_run.py. Run under gpurun; logs -> gpurun_out/sanitize_*.log,
# summaries are committed under profiles/.
#   tools/sanitize.sh [steps]
set -u
steps=${1:-50}
mkdir -p gpurun_out
rc=0
python tools/sanitize_run.py 3 > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
for tool in memcheck racecheck synccheck initcheck; do
    s=$steps
    [ "$tool" = racecheck ] && s=$(( steps < 12 ? steps : 12 ))     # racecheck slows the kernel ~100x
    [ "$tool" = initcheck ] && s=$(( steps < 12 ? steps : 12 ))
    compute-sanitizer --tool "$tool" --print-limit 20 python tools/sanitize_run.py "$s" > "gpurun_out/sanitize_$tool.log" 2>&1
    code=$?
    summary=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "gpurun_out/sanitize_$tool.log" | tail -1)
    echo "$tool: exit $code; ${summary:-no summary line}; $(grep -c 'sanitize_run ok' gpurun_out/sanitize_$tool.log) completed run(s)"
    [ $code -ne 0 ] && rc=1
done
exit $rc
