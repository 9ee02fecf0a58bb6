#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tune3.log
for LIBV in default SHF ALUSUB BOTH; do
  if [ "$LIBV" == "default" ]; then unset MCX_B200_LIB; else export MCX_B200_LIB=$PWD/montecarlox.jl_b200/lib/variants/libmcx_$LIBV.so; fi
  for rep in 1 2; do
    out=$(timeout 300 python bench.py --steps 2 --warmup 2 --sweeps-per-step 20 --no-pt --no-cpu 2>&1 | tail -1)
    echo "LIB=$LIBV $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value=%.1f kernel=%.1f frac=%.3f e2e=%.1f" % (d["value"], d["roofline"]["kernel_attempts_per_ns"], d["roofline"]["frac"], d["e2e"]["value"]))' 2>&1)" | tee -a gpurun_out/tune3.log
  done
done
