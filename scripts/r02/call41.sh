#!/bin/bash
# round 2, GPU call 41: graph-replayed PT rounds on one rank: chain groups vs one launch per half-sweep, rounds per replay
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call41.log
: > $O
for env in "MCX_NOP=1" "MCX_GROUPS=0" "MCX_PT_GRAPH=32" "MCX_PT_GRAPH=32 MCX_GROUPS=0"; do
  echo "== $env (MCX_PT_PERSIST=0, every 1)" >> $O
  env $env MCX_PT_PERSIST=0 timeout 300 python scripts/bench_pt_rank.py --counts 256,64,32 --every 1 --rounds 640 2>&1 | python -c "
import sys, re
for l in sys.stdin:
    if l.startswith('{'):
        g = lambda k: re.search(r'\"%s\": ([^,}]+)' % k, l).group(1)
        print('  %3s replicas: %9.0f sweeps/s %7.1f attempts/ns host %s us/sweep [%s]' % (g('replicas_on_rank'), float(g('rank_sweeps_per_s')), float(g('attempts_per_ns')), g('host_enqueue_us_per_sweep'), g('path')))" >> $O
done
cat $O
