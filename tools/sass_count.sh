#!/bin/bash
# usage: tools/sass_count.sh lib.so  -> number of SASS instructions of wt_step_kernel
cuobjdump -sass "$1" | awk '/Function :/{f=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/{c[f]++} END{for(k in c) if (k ~ /step_kernel/) print k, c[k]}'
