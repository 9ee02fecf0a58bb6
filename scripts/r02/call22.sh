#!/bin/bash
# round 2, GPU call 22: k_ising2d with per-thread cp.async staging of the next trip (MCX_STAGE=1) against direct loads
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call22.log
: > $O
( MCX_STAGE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_oracle.py tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r02/call22_pytest.log 2>&1
tail -3 gpurun_out/r02/call22_pytest.log
for rep in 1 2; do for st in 0 1; do
  MCX_STAGE=$st python bench.py --no-cpu --no-pt --no-extras --steps 3 --warmup 3 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('STAGE=$st value=%.1f kernel=%.1f frac=%.3f' % (d['value'], d['roofline']['kernel_attempts_per_ns'], d['roofline']['frac']))" >> $O
done; done
cat $O
