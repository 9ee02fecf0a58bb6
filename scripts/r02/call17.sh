#!/bin/bash
# round 2, GPU call 17 (4 GPUs): parallel tempering with the rounds queued by the host against the persistent launch, across ranks
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call17.log
: > $O
for mode in 0 1; do
  echo "== --config c3 N=4 MCX_PT_PERSIST=$mode" >> $O
  MCX_PT_PERSIST=$mode timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2984$mode bench.py --gpus 4 --config c3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('every 200: %.1f  every sweep: %.1f  sha %s %s' % (d['value'], d['every_sweep']['value'], d['parity']['labels_and_energies_sha'], d['every_sweep']['parity']['labels_and_energies_sha']))" >> $O 2>&1
done
echo "== windows with neighbour exchanges over NCCL, N=4" >> $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29852 scripts/nccl_windows_check.py 2>&1 | grep "^{" >> $O
cat $O
