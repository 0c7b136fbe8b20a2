#!/bin/bash
# usage: tools/run_variants.sh out.log name=lib.so[:ENV=VAL...] ...   (runs tools/prof_step.py per variant)
out=$1; shift
: > $out
for spec in "$@"; do
  name=${spec%%=*}; rest=${spec#*=}
  lib=${rest%%:*}; envs=""
  if [[ "$rest" == *:* ]]; then envs=$(echo "${rest#*:}" | tr ':' ' '); fi
  echo "== $name ($lib $envs)" >> $out
  env WT_B200_LIB=$PWD/$lib $envs timeout 300 python tools/prof_step.py --plants ${PLANTS:-262144} --zones ${ZONES:-10} --steps 3 --warmup 2 ${EXTRA} >> $out 2>&1
done
cat $out
