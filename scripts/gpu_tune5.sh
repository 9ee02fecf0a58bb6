#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tune5.log
run() { out=$(timeout 300 python bench.py --steps 2 --warmup 2 --sweeps-per-step 20 --no-pt --no-cpu 2>&1 | tail -1); echo "$1 $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value=%.1f kernel=%.1f frac=%.3f" % (d["value"], d["roofline"]["kernel_attempts_per_ns"], d["roofline"]["frac"]))' 2>&1)" | tee -a gpurun_out/tune5.log; }
export MCX_B200_LIB=$PWD/montecarlox.jl_b200/lib/variants/libmcx_NOP.so
for V in 3 12 11; do export MCX_VARIANT=$V; run "NOPHILOX+SEQROWS V=$V"; done
