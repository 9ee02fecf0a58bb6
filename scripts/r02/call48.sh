#!/bin/bash
# round 2, GPU call 48 (2 GPUs): the driver's own N = 2 command on the final tree
mkdir -p gpurun_out/r02
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r02/call48_bench_n2.json 2> gpurun_out/r02/call48_bench_n2.err
grep real gpurun_out/r02/call48_bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02/call48_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value=%.1f frac=%.3f e2e=%.1f (bit %.1f) pt=%.0f pt_every=%.0f slab_strong=%.0f slab_weak=%.0f' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['bit_buffers']['value'], d['pt']['value'], d['pt_every_sweep']['value'], d['slab_strong']['value'], d['slab_weak']['value']))
print('   sha', d['pt']['parity']['labels_and_energies_sha'], d['pt_every_sweep']['parity']['labels_and_energies_sha'], d['slab_strong']['parity'])
print('   configs', {k: (round(v['value'], 2) if isinstance(v, dict) and 'value' in v else v) for k, v in d.get('configs', {}).items()})
PY
