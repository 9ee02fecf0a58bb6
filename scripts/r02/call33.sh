#!/bin/bash
# round 2, GPU call 33 (2 GPUs): the reworked strip loop on the multi-GPU paths: parity on distinct devices (slabs over IPC, PT
# peers, windows, histogram all-reduce), then the bench under torchrun at N = 2 and the N = 1 line on the same box
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call33.log
: > $O
nvidia-smi -L >> $O 2>&1
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slab.py -x -q 2>&1 | tail -6 ) > gpurun_out/r02/call33_pytest.log 2>&1
tail -4 gpurun_out/r02/call33_pytest.log
echo "== bench N=2" >> $O
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r02/call33_bench_n2.json 2> gpurun_out/r02/call33_bench_n2.err
tail -4 gpurun_out/r02/call33_bench_n2.err | grep -v OMP >> $O
echo "== bench N=1" >> $O
( time timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu ) > gpurun_out/r02/call33_bench_n1.json 2> gpurun_out/r02/call33_bench_n1.err
tail -4 gpurun_out/r02/call33_bench_n1.err >> $O
python - <<'PY' >> gpurun_out/r02/call33.log
import json
for n in (1, 2):
    try:
        d = json.loads(open('gpurun_out/r02/call33_bench_n%d.json' % n).read().strip().splitlines()[-1])
        print('N=%d value=%.1f frac=%.3f e2e=%.1f pt=%.0f pt_every=%.0f slab_strong=%s slab_weak=%s' % (n, d['value'], d['roofline']['frac'], d['e2e']['value'], d['pt']['value'], d['pt_every_sweep']['value'], (d.get('slab_strong') or {}).get('value'), (d.get('slab_weak') or {}).get('value')))
        print('   parity', {k: v.get('parity') for k, v in d.get('configs', {}).items() if isinstance(v, dict) and v.get('parity')}, d['pt'].get('parity'), d['pt_every_sweep'].get('parity'))
    except Exception as e:
        print('N=%d: %r' % (n, e))
PY
cut -c1-300 $O
