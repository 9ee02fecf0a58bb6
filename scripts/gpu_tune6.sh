#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tune6.log
run() { out=$(timeout 300 python bench.py --steps 2 --warmup 2 --sweeps-per-step 20 --no-cpu 2>&1 | tail -1); echo "$1 $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value=%.1f kernel=%.1f frac=%.3f pt=%.0f (%.0f/ns) pt1=%.0f" % (d["value"], d["roofline"]["kernel_attempts_per_ns"], d["roofline"]["frac"], d["pt"]["value"], d["pt"]["attempts_per_ns"], d["pt_every_sweep"]["value"]))' 2>&1)" | tee -a gpurun_out/tune6.log; }
export MCX_NO_GRID_TRIM=1; run "untrimmed grid"
unset MCX_NO_GRID_TRIM; run "trimmed grid"; run "trimmed grid"
