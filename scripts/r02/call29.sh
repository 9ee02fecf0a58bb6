#!/bin/bash
# round 2, GPU call 29: row bands ordered by flags inside the kernel (default) against CUDA events (MCX_BAND_FLAGS=0)
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call29.log
: > $O
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_oracle.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r02/call29_pytest.log 2>&1
tail -2 gpurun_out/r02/call29_pytest.log
for rep in 1 2; do for fl in 0 1; do
  MCX_BAND_FLAGS=$fl timeout 300 python bench.py --no-cpu --no-pt --no-extras --steps 3 --warmup 3 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('FLAGS=$fl value=%.1f kernel=%.1f frac=%.3f e2e=%.1f' % (d['value'], d['roofline']['kernel_attempts_per_ns'], d['roofline']['frac'], d['e2e']['value']))" >> $O
done; done
for sh in "MCX_BANDS=8 MCX_BAND_ROWS=16" "MCX_BANDS=8 MCX_BAND_ROWS=32" "MCX_BANDS=16 MCX_BAND_ROWS=16" "MCX_BANDS=4 MCX_BAND_ROWS=32"; do
  env $sh timeout 300 python bench.py --no-cpu --no-pt --no-extras --steps 3 --warmup 3 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$sh value=%.1f kernel=%.1f' % (d['value'], d['roofline']['kernel_attempts_per_ns']))" >> $O
done
cat $O
