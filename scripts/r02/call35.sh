#!/bin/bash
# round 2, GPU call 35 (2 GPUs): graph-replayed parallel-tempering rounds across ranks (peer stores): parity digests and rates
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call35.log
: > $O
( timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pt_persistent.py -x -q 2>&1 | tail -4 ) > gpurun_out/r02/call35_pytest.log 2>&1
tail -2 gpurun_out/r02/call35_pytest.log
for g in 0 1; do
  echo "== --config c3 N=2 MCX_PT_GRAPH=$g" >> $O
  MCX_PT_GRAPH=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2985$g bench.py --gpus 2 --config c3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('every 200: %.1f  every sweep: %.1f  sha %s %s' % (d['value'], d['every_sweep']['value'], d.get('parity', {}).get('labels_and_energies_sha'), d['every_sweep'].get('parity', {}).get('labels_and_energies_sha')))" >> $O
done
cat $O
