#!/bin/bash
# round 2, GPU call 50: graph-replay tests after the bookkeeping clean-up of the sweep-series graph
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pt_persistent.py -m gpu -x -q -k "graph" 2>&1 | tail -2
timeout 60 python bench.py --no-cpu --no-pt --no-extras --steps 2 --warmup 2 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value=%.1f frac=%.3f e2e=%.1f energy=%s' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['result']['energy_per_site']))"
