#!/bin/bash
# round 2, GPU call 36 (8 GPUs): the round's bench lines at N = 8 and N = 4 after the strip-loop rework and with the parallel-tempering
# rounds replayed from CUDA graphs; config 3 at N = 8 with the graph replay switched off for comparison
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call36.log
: > $O
nvidia-smi -L | wc -l >> $O 2>&1
echo "== bench N=8" >> $O
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r02/call36_bench_n8.json 2> gpurun_out/r02/call36_bench_n8.err
grep real gpurun_out/r02/call36_bench_n8.err >> $O
echo "== bench --config c3 N=8 MCX_PT_GRAPH=0" >> $O
( time MCX_PT_GRAPH=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29833 bench.py --gpus 8 --config c3 --no-cpu ) > gpurun_out/r02/call36_bench_c3_n8_nograph.json 2> gpurun_out/r02/call36_c3.err
grep real gpurun_out/r02/call36_c3.err >> $O
echo "== bench N=4" >> $O
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29834 bench.py --gpus 4 --steps 3 --warmup 3 ) > gpurun_out/r02/call36_bench_n4.json 2> gpurun_out/r02/call36_bench_n4.err
grep real gpurun_out/r02/call36_bench_n4.err >> $O
python - <<'PY' >> gpurun_out/r02/call36.log
import json
for n, f in ((8, 'call36_bench_n8.json'), (4, 'call36_bench_n4.json')):
    try:
        d = json.loads(open('gpurun_out/r02/' + f).read().strip().splitlines()[-1])
        print('N=%d value=%.1f frac=%.3f e2e=%.1f pt=%.0f pt_every=%.0f slab_strong=%s slab_weak=%s' % (n, d['value'], d['roofline']['frac'], d['e2e']['value'], d['pt']['value'], d['pt_every_sweep']['value'], (d.get('slab_strong') or {}).get('value'), (d.get('slab_weak') or {}).get('value')))
        print('   sha', d['pt'].get('parity', {}).get('labels_and_energies_sha'), d['pt_every_sweep'].get('parity', {}).get('labels_and_energies_sha'), (d.get('slab_strong') or {}).get('parity'))
    except Exception as e:
        print('N=%d: %r' % (n, e))
try:
    d = json.loads(open('gpurun_out/r02/call36_bench_c3_n8_nograph.json').read().strip().splitlines()[-1])
    print('c3 N=8 no graph: every 200: %.1f  every sweep: %.1f' % (d['value'], d['every_sweep']['value']))
except Exception as e:
    print('c3: %r' % e)
PY
cut -c1-300 $O
