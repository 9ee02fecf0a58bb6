#!/bin/bash
# round 2, GPU call 45: final tree: parity of the streaming paths, then the default bench line
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call45.log
: > $O
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_oracle.py tests/test_gpu_full_size.py tests/test_gpu_slab.py tests/test_gpu_checkpoint.py tests/test_gpu_statistics.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r02/call45_pytest.log 2>&1
tail -2 gpurun_out/r02/call45_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/r02/call45_bench.json 2> gpurun_out/r02/call45_bench.err
python - <<'PY' >> gpurun_out/r02/call45.log
import json
d = json.loads(open('gpurun_out/r02/call45_bench.json').read().strip().splitlines()[-1])
print('value=%.1f frac=%.3f kernel=%.1f kernel_ms=%.5f e2e=%.1f (serial %.1f bit %.1f) pt=%.0f pt_every=%.0f launches=%s' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_attempts_per_ns'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['serial']['value'], d['e2e']['bit_buffers']['value'], d['pt']['value'], d['pt_every_sweep']['value'], d['gpu_launches']))
print({k: (round(v['value'], 2) if isinstance(v, dict) and 'value' in v else v) for k, v in d.get('configs', {}).items()})
PY
cat $O
