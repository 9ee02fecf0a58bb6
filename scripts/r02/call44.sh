#!/bin/bash
# round 2, GPU call 44: two trips per loop iteration (probe build u2) against the default strip loop
mkdir -p gpurun_out/r02
bash scripts/gpu_ab.sh default u2 > gpurun_out/r02/call44.log 2>&1
for v in default u2; do
  if [ "$v" = default ]; then unset MCX_B200_LIB; else export MCX_B200_LIB=$PWD/montecarlox.jl_b200/lib/libmcx_b200_$v.so; fi
  timeout 200 python scripts/bench_pt_rank.py --counts 256,32 --every 200 --rounds 3 2>&1 | python -c "
import sys, re
for l in sys.stdin:
    if l.startswith('{'):
        g = lambda k: re.search(r'\"%s\": ([^,}]+)' % k, l).group(1)
        print('PT LIB=$v %3s replicas: %9.0f sweeps/s %7.1f attempts/ns' % (g('replicas_on_rank'), float(g('rank_sweeps_per_s')), float(g('attempts_per_ns'))))" >> gpurun_out/r02/call44.log
done
cat gpurun_out/r02/call44.log
