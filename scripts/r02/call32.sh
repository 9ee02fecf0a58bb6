#!/bin/bash
# round 2, GPU call 32: end-to-end rate (int8 host buffers, H2D inside the timed region) against the number of launches per half-sweep
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call32.log
: > $O
for sh in "MCX_NOP=1" "MCX_BANDS=8 MCX_BAND_ROWS=32" "MCX_BANDS=4 MCX_BAND_ROWS=32" "MCX_BANDS=8 MCX_BAND_ROWS=16" "MCX_NOP=1"; do
  env $sh timeout 300 python bench.py --no-cpu --no-pt --no-extras --steps 5 --warmup 3 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
e = d['e2e']
print('$sh value=%.1f kernel=%.1f e2e=%.1f (%.2f ms) serial=%.1f bit_buffers=%.1f with_download=%.1f' % (d['value'], d['roofline']['kernel_attempts_per_ns'], e['value'], e['ms_per_step'], e['serial']['value'], e['bit_buffers']['value'], e['bit_buffers_with_download']['value']))" >> $O
done
cat $O
