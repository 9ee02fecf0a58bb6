#!/bin/bash
# round 2, GPU call 42: full GPU suite, smoke, the default bench line and the reference arm on the final tree
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call42.log
: > $O
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) > gpurun_out/r02/call42_pytest.log 2>&1
tail -6 gpurun_out/r02/call42_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $O 2>&1
echo "== bench default" >> $O
( time timeout 900 python bench.py ) > gpurun_out/r02/call42_bench.json 2> gpurun_out/r02/call42_bench.err
tail -4 gpurun_out/r02/call42_bench.err >> $O
echo "== bench --impl reference" >> $O
( time timeout 400 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02/call42_bench_ref.json 2>> $O
python - <<'PY' >> gpurun_out/r02/call42.log
import json
d = json.loads(open('gpurun_out/r02/call42_bench.json').read().strip().splitlines()[-1])
print('value=%.1f frac=%.3f kernel=%.1f e2e=%.1f (serial %.1f bit %.1f) pt=%.0f pt_every=%.0f launches=%s clocks=%s' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_attempts_per_ns'], d['e2e']['value'], d['e2e']['serial']['value'], d['e2e']['bit_buffers']['value'], d['pt']['value'], d['pt_every_sweep']['value'], d['gpu_launches'], d['clocks']))
print({k: (round(v['value'], 2) if isinstance(v, dict) and 'value' in v else v) for k, v in d.get('configs', {}).items()})
print('cpu_baseline', d.get('cpu_baseline'))
r = json.loads(open('gpurun_out/r02/call42_bench_ref.json').read().strip().splitlines()[-1])
print('reference', r.get('value'), r.get('cpu_baseline'))
PY
cut -c1-400 $O
