#!/bin/bash
# usage (8-GPU box): tools/scale_run.sh  -> gpurun_out/r2_bench_n{1,2,4,8}.json (+ N=8 without deferral)
mkdir -p gpurun_out
run() { n=$1; shift; out=$1; shift
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline "$@" > $out 2> gpurun_out/scale_err_$n.log
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 5 "$@" > $out 2> gpurun_out/scale_err_$n.log; fi
  echo "N=$n rc=$? $(python -c "import json,sys; d=json.loads([l for l in open('$out') if l.startswith('{')][-1]); print('value %.4e ms/step %.3f e2e %.4e frac %.4f alone %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_alone']['frac']))" 2>&1)"; }
run 8 gpurun_out/r2_bench_n8.json
run 4 gpurun_out/r2_bench_n4.json
run 2 gpurun_out/r2_bench_n2.json
run 1 gpurun_out/r2_bench_n1_8gpu_box.json
run 8 gpurun_out/r2_bench_n8_nodeferral.json --catch-up-floor-div 0
