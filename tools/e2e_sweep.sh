#!/bin/bash
# usage (GPU box): tools/e2e_sweep.sh  -> ms per wt_step_host call for several slab plans, 1M and 131,072 plants
for P in 1048576 131072; do
  echo "== PLANTS=$P"
  PLANTS=$P python tools/e2e_probe.py
  for mx in 65536 262144; do PLANTS=$P WT_B200_HOST_SLAB_MAX=$mx python tools/e2e_probe.py | sed "s/^/max=$mx /"; done
  for mn in 4096 32768; do PLANTS=$P WT_B200_HOST_SLAB_MIN=$mn python tools/e2e_probe.py | sed "s/^/min=$mn /"; done
  for k in 8 16 32; do PLANTS=$P WT_B200_HOST_SLABS=$k python tools/e2e_probe.py; done
done
