#!/bin/bash
# round 2, GPU call 39 (8 GPUs): the default bench line at N = 8 with sweep series and PT rounds replayed from CUDA graphs
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call39.log
: > $O
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r02/call39_bench_n8.json 2> gpurun_out/r02/call39_bench_n8.err
grep real gpurun_out/r02/call39_bench_n8.err >> $O
python - <<'PY' >> gpurun_out/r02/call39.log
import json
d = json.loads(open('gpurun_out/r02/call39_bench_n8.json').read().strip().splitlines()[-1])
print('N=8 value=%.1f frac=%.3f e2e=%.1f (serial %.1f, bit_buffers %.1f) pt=%.0f pt_every=%.0f slab_strong=%s slab_weak=%s' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['serial']['value'], d['e2e']['bit_buffers']['value'], d['pt']['value'], d['pt_every_sweep']['value'], (d.get('slab_strong') or {}).get('value'), (d.get('slab_weak') or {}).get('value')))
print('   sha', d['pt'].get('parity', {}).get('labels_and_energies_sha'), d['pt_every_sweep'].get('parity', {}).get('labels_and_energies_sha'), (d.get('slab_strong') or {}).get('parity'))
print('   configs', {k: (round(v['value'], 2) if isinstance(v, dict) and 'value' in v else None) for k, v in d.get('configs', {}).items()})
PY
cut -c1-400 $O
