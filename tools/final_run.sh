#!/bin/bash
# usage (1-GPU box): tools/final_run.sh -> the round's evidence under gpurun_out/ (tests, bench lines, launch list)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s --durations=8 > gpurun_out/r2_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_gpu_tests.log
python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/final_err.log; echo "reference arm rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2>> gpurun_out/final_err.log; echo "bench rc=$?"
python bench.py --steps 1900 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n1_1900steps_floor8.json 2>> gpurun_out/final_err.log; echo "1900 rc=$?"
python bench.py --steps 1900 --warmup 5 --no-cpu-baseline --catch-up-floor-div 0 > gpurun_out/r2_bench_n1_1900steps_halt.json 2>> gpurun_out/final_err.log; echo "1900 halt rc=$?"
bash tools/launch_list.sh gpurun_out/r2_launch_list_bench.csv
