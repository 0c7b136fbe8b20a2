#!/bin/bash
# usage (on the GPU box): tools/gpu_variants.sh v1 v2 ...   -> natural-order and cost-sorted step times of tools/variants/<v>.so
for v in "$@"; do
  echo "== $v"; WT_B200_LIB=$PWD/tools/variants/$v.so python tools/prof_step.py --plants 262144 --zones 10 --steps 3 --warmup 2 2>&1 | tail -3 | head -1
  echo "== $v sorted"; WT_B200_LIB=$PWD/tools/variants/$v.so python tools/prof_step.py --plants 262144 --zones 10 --steps 6 --warmup 4 --sort-every 2 2>&1 | tail -3 | head -1
done
