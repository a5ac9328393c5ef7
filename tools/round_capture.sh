#!/bin/bash
# Everything profiles/ holds for a round, under gpurun (one GPU): bench lines of every config, the reference arm, the ncu
# launch list of the bench command, ncu --set full captures of the step kernel, the parity report, kprobe tables.
#   tools/round_capture.sh r2b bench     # bench lines, launch list, parity report, kprobe (small files)
#   tools/round_capture.sh r2b ncu       # ncu --set full: fixed_hip fresh + steady (2 x 22 MB)
#   tools/round_capture.sh r2b ncu5      # ncu --set full: free_hip steady
# (gpurun brings back at most 64 MiB per call: the three captures do not fit one call)
r=${1:-r2}
part=${2:-bench}
mkdir -p gpurun_out
if [ "$part" = bench ]; then
python bench.py > gpurun_out/${r}_bench.json 2> gpurun_out/${r}_bench.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${r}_bench_driver_shape.json 2>> gpurun_out/${r}_bench.err
python bench.py --config 4 --steps 500 --no-cpu-baseline > gpurun_out/${r}_bench_config4.json 2>> gpurun_out/${r}_bench.err
python bench.py --config 5 --steps 500 --no-cpu-baseline > gpurun_out/${r}_bench_config5.json 2>> gpurun_out/${r}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${r}_bench_reference.json 2>> gpurun_out/${r}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${r}_launches.csv \
    python bench.py --steps 100 --warmup 10 --preroll 300 --no-cpu-baseline > gpurun_out/${r}_launches.log 2>&1
python tools/parity_report.py > gpurun_out/${r}_parity_stdout.log 2>&1
python tools/kprobe.py std cfg4 small modes > gpurun_out/${r}_kprobe_stdout.log 2>&1
tail -3 gpurun_out/${r}_bench.err
elif [ "$part" = ncu ]; then
tools/ncu_capture.sh ${r}_fresh 8 fixed_hip
tools/ncu_capture.sh ${r}_steady 1500 fixed_hip
else
tools/ncu_capture.sh ${r}_free_hip_steady 1500 free_hip
fi
