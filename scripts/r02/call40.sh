#!/bin/bash
# round 2, GPU call 40 (2 GPUs): the fused tail of a graph-replayed PT round (publish + wait + exchange + clock in one launch)
# against the three launches (MCX_PT_GRAPH=2): parity, digests, rates; single-rank rates at 32 / 64 replicas
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call40.log
: > $O
( timeout 400 python -m pytest tests/test_gpu_pt_persistent.py -x -q 2>&1 | tail -3 ) > gpurun_out/r02/call40_pytest.log 2>&1
tail -2 gpurun_out/r02/call40_pytest.log
for g in 2 1; do
  echo "== --config c3 N=2 MCX_PT_GRAPH=$g" >> $O
  MCX_PT_GRAPH=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2985$g bench.py --gpus 2 --config c3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('every 200: %.1f  every sweep: %.1f  sha %s %s' % (d['value'], d['every_sweep']['value'], d.get('parity', {}).get('labels_and_energies_sha'), d['every_sweep'].get('parity', {}).get('labels_and_energies_sha')))" >> $O
  echo "== one rank, MCX_PT_GRAPH=$g MCX_PT_PERSIST=0, every 1" >> $O
  MCX_PT_GRAPH=$g MCX_PT_PERSIST=0 timeout 300 python scripts/bench_pt_rank.py --counts 64,32 --every 1 --rounds 600 2>&1 | python -c "
import sys, re
for l in sys.stdin:
    if l.startswith('{'):
        g = lambda k: re.search(r'\"%s\": ([^,}]+)' % k, l).group(1)
        print('  %3s replicas: %9.0f sweeps/s %7.1f attempts/ns host %s us/sweep [%s]' % (g('replicas_on_rank'), float(g('rank_sweeps_per_s')), float(g('attempts_per_ns')), g('host_enqueue_us_per_sweep'), g('path')))" >> $O
done
cat $O
