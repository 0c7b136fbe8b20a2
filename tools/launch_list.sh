#!/bin/bash
# usage (GPU box): tools/launch_list.sh <out.csv>   -> per-launch device times of a short bench run (cold-cache, serialised: compare SHARES)
out=${1:-gpurun_out/launches.csv}
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/launch_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $out python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/launch_ncu.log 2>&1
tail -2 gpurun_out/launch_ncu.log
