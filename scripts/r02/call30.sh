#!/bin/bash
# round 2, GPU call 30: one-bit planes on the one-block-per-trip loop (parity, rate); config-1 leg (resident kernel) against
# the previous round's library
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call30.log
: > $O
( timeout 600 python -m pytest tests/test_gpu_bits.py tests/test_gpu_full_oracle.py tests/test_gpu_resident.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r02/call30_pytest.log 2>&1
tail -2 gpurun_out/r02/call30_pytest.log
for v in base default; do
  if [ "$v" = default ]; then unset MCX_B200_LIB; else export MCX_B200_LIB=$PWD/montecarlox.jl_b200/lib/libmcx_b200_$v.so; fi
  for rep in 1 2; do
    timeout 300 python bench.py --storage bit --no-cpu --no-pt --no-extras --steps 3 --warmup 3 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('bit LIB=$v value=%.1f kernel=%.1f' % (d['value'], d['roofline']['kernel_attempts_per_ns']))" >> $O
    timeout 300 python bench.py --config c1 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c1 LIB=$v value=%.3f' % d['value'])" >> $O
  done
done
cat $O
