#!/bin/bash
# usage: tools/ncu_step.sh <lib.so> <out-name> [kernel-regex]  -> gpurun_out/<out-name>.ncu-rep (one launch of the step's kernels, 262144 x 10)
lib=$1; name=$2; pat=${3:-wt_step_run_kernel}
mkdir -p gpurun_out
WT_B200_LIB=$PWD/$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 \
  -f -o gpurun_out/$name python tools/prof_step.py --plants ${PLANTS:-262144} --zones ${ZONES:-10} --steps 1 --warmup 2 > gpurun_out/$name.log 2>&1
tail -3 gpurun_out/$name.log
