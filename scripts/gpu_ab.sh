#!/bin/bash
# A/B of library variants on ONE box (boxes differ by several per cent, so never compare across calls):
#   scripts/gpu_ab.sh default queue pfl1 ...   (names of lib/libmcx_b200_NAME.so; "default" = the shipped library)
cd "$(dirname "$0")/.."
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = default ]; then unset MCX_B200_LIB; else export MCX_B200_LIB=$PWD/montecarlox.jl_b200/lib/libmcx_b200_$v.so; fi
  python bench.py --no-cpu --no-pt --no-extras --steps 3 --warmup 3 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('LIB=%-10s value=%.1f kernel=%.1f frac=%.3f' % ('$v', d['value'], d['roofline']['kernel_attempts_per_ns'], d['roofline']['frac']))"
done
done
