#!/bin/bash
# round 2, GPU call 26: L2 prefetch of the next trip's rows (probe build) against the default; config-1 leg (resident kernel) against the previous round's library
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call26.log
: > $O
bash scripts/gpu_ab.sh default pf >> $O 2>&1
for v in base default; do
  if [ "$v" = default ]; then unset MCX_B200_LIB; else export MCX_B200_LIB=$PWD/montecarlox.jl_b200/lib/libmcx_b200_$v.so; fi
  for rep in 1 2; do
  python bench.py --config c1 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c1 LIB=$v value=%.3f ms_per_step=%.3f' % (d['value'], d['ms_per_step']))" >> $O
  done
done
cat $O
