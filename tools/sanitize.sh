#!/bin/bash
# In-kernel checks in place of compute-sanitizer (which is closed on this GPU pool: "runs under it have left GPUs
# needing a reset"): builds the CHECKED library (-DOS2R_CHECKED: lane-sort permutation, env-index bounds, terminal-record
# capacity and shared-memory guard words validated inside the step kernel, violations counted) and runs
# tools/sanitize_run.py against it TWICE: all counters must be zero and the two runs' final states bit-identical
# (a shared-memory race between the sort scratch and the cold slots, or across the per-iteration barriers, would make
# the result depend on warp timing). Run under gpurun; log -> gpurun_out/sanitize_checked.log (summary under profiles/).
#   tools/sanitize.sh [steps]
set -u
steps=${1:-50}
mkdir -p gpurun_out
lib=gym_os2r_b200/csrc/libos2r_checked.so
make -C gym_os2r_b200/csrc libos2r_checked.so > /dev/null || exit 1     # incremental: rebuilt only when the sources changed
log=gpurun_out/sanitize_checked.log
: > "$log"
for run in 1 2; do
    echo "== run $run" >> "$log"
    OS2R_LIB="$PWD/$lib" python tools/sanitize_run.py "$steps" >> "$log" 2>&1 || { echo "run $run failed"; tail -5 "$log"; exit 1; }
done
grep -E "check\[|library|state digest|sanitize_run" "$log"
[ "$(grep -c 'sanitize_run ok' "$log")" = 2 ] || exit 1
[ "$(grep 'state digest' "$log" | sort -u | wc -l)" = 1 ] || { echo "the two runs differ: nondeterministic"; exit 1; }
echo "sanitize: all checks zero, two runs bit-identical"
