#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tune4.log
for PF in 0 1 2 3 4; do
 for R in 16 32 64; do
  export MCX_L2_PREFETCH_TRIPS=$PF MCX_ROWS_PER_STRIP=$R
  out=$(timeout 300 python bench.py --steps 2 --warmup 2 --sweeps-per-step 20 --no-pt --no-cpu 2>&1 | tail -1)
  echo "PF=$PF R=$R $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value=%.1f kernel=%.1f frac=%.3f" % (d["value"], d["roofline"]["kernel_attempts_per_ns"], d["roofline"]["frac"]))' 2>&1)" | tee -a gpurun_out/tune4.log
 done
done
