#!/bin/bash
# round 2, GPU call 38: sweep-series graph at 32 sweeps per replay (default) against none
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call38.log
: > $O
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graph or bands" 2>&1 | tail -3 ) > gpurun_out/r02/call38_pytest.log 2>&1
tail -2 gpurun_out/r02/call38_pytest.log
for rep in 1 2; do for g in 0 1 16; do
  MCX_SWEEP_GRAPH=$g timeout 300 python bench.py --no-cpu --no-pt --no-extras --steps 5 --warmup 3 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
e = d['e2e']
print('SWEEP_GRAPH=$g value=%.1f kernel=%.1f frac=%.3f e2e=%.1f (%.2f ms) serial=%.1f bit_buffers=%.1f' % (d['value'], d['roofline']['kernel_attempts_per_ns'], d['roofline']['frac'], e['value'], e['ms_per_step'], e['serial']['value'], e['bit_buffers']['value']))" >> $O
done; done
cat $O
