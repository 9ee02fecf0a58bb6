#!/bin/bash
# round 2, GPU call 37: series of row-band sweeps replayed from a CUDA graph: parity, then the bench line with and without it
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call37.log
: > $O
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_oracle.py tests/test_gpu_full_size.py tests/test_gpu_checkpoint.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r02/call37_pytest.log 2>&1
tail -3 gpurun_out/r02/call37_pytest.log
for rep in 1 2; do for g in 0 1; do
  MCX_SWEEP_GRAPH=$g timeout 300 python bench.py --no-cpu --no-pt --no-extras --steps 5 --warmup 3 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
e = d['e2e']
print('SWEEP_GRAPH=$g value=%.1f kernel=%.1f frac=%.3f e2e=%.1f (%.2f ms) serial=%.1f bit_buffers=%.1f launches=%s' % (d['value'], d['roofline']['kernel_attempts_per_ns'], d['roofline']['frac'], e['value'], e['ms_per_step'], e['serial']['value'], e['bit_buffers']['value'], d['gpu_launches']))" >> $O
done; done
cat $O
