#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tune2.log
for V in 0 3 8; do
  for R in 16; do
    export MCX_VARIANT=$V MCX_ROWS_PER_STRIP=$R
    out=$(timeout 300 python bench.py --steps 2 --warmup 2 --sweeps-per-step 20 --no-pt --no-cpu --track ${TRACK:-0} 2>&1 | tail -1)
    echo "V=$V R=$R $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value=%.1f kernel=%.1f frac=%.3f e2e=%.1f" % (d["value"], d["roofline"]["kernel_attempts_per_ns"], d["roofline"]["frac"], d["e2e"]["value"]))' 2>&1)" | tee -a gpurun_out/tune2.log
  done
done
