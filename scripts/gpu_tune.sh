#!/bin/bash
# Sweep the strip height / residency knobs of the fast kernel; prints kernel-only attempts/ns per setting.
mkdir -p gpurun_out
: > gpurun_out/tune.log
for R in 4 8 16 32 64; do
  for C in 0 4 5; do
    if [ "$C" == "0" ]; then unset MCX_CTAS_PER_SM; else export MCX_CTAS_PER_SM=$C; fi
    export MCX_ROWS_PER_STRIP=$R
    out=$(timeout 300 python bench.py --steps 2 --warmup 2 --sweeps-per-step 20 --no-pt --no-cpu --track ${TRACK:-0} 2>&1 | tail -1)
    echo "R=$R CTAS=$C $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value=%.1f kernel=%.1f frac=%.3f e2e=%.1f" % (d["value"], d["roofline"]["kernel_attempts_per_ns"], d["roofline"]["frac"], d["e2e"]["value"]))' 2>&1)" | tee -a gpurun_out/tune.log
  done
done
