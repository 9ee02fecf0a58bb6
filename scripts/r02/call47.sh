#!/bin/bash
# round 2, GPU call 47: two trips per loop iteration in the one-bit kernel (probe build bu2): parity and rate
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call47.log
: > $O
( MCX_B200_LIB=$PWD/montecarlox.jl_b200/lib/libmcx_b200_bu2.so timeout 200 python -m pytest tests/test_gpu_bits.py -m gpu -x -q 2>&1 | tail -2 ) >> $O 2>&1
for rep in 1 2; do for v in default bu2; do
  if [ "$v" = default ]; then unset MCX_B200_LIB; else export MCX_B200_LIB=$PWD/montecarlox.jl_b200/lib/libmcx_b200_$v.so; fi
  timeout 200 python bench.py --storage bit --no-cpu --no-pt --no-extras --steps 3 --warmup 3 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('bit LIB=$v value=%.1f kernel=%.1f' % (d['value'], d['roofline']['kernel_attempts_per_ns']))" >> $O
done; done
cat $O
