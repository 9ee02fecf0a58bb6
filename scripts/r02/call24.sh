#!/bin/bash
# round 2, GPU call 24: the shared strip loop (k_strip.cuh) in the ticket-queue and persistent-rounds kernels: parity, then
# the per-rank parallel-tempering rates at 5 / 6 / 8 CTAs per SM against the library of the previous round
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call24.log
: > $O
( time timeout 900 python -m pytest tests/test_gpu_queue.py tests/test_gpu_pt_persistent.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r02/call24_pytest.log 2>&1
tail -4 gpurun_out/r02/call24_pytest.log
for v in base default q6 q8; do
  if [ "$v" = default ]; then unset MCX_B200_LIB; else export MCX_B200_LIB=$PWD/montecarlox.jl_b200/lib/libmcx_b200_$v.so; fi
  echo "== LIB=$v every 200" >> $O
  timeout 300 python scripts/bench_pt_rank.py --counts 256,64,32 --every 200 --rounds 3 2>&1 | cut -c1-250 >> $O
  echo "== LIB=$v every 1" >> $O
  timeout 300 python scripts/bench_pt_rank.py --counts 64,32 --every 1 --rounds 600 2>&1 | cut -c1-250 >> $O
done
python - <<'PY'
import json
for l in open('gpurun_out/r02/call24.log'):
    if l.startswith('=='): print(l.strip())
    elif l.startswith('{'):
        d=json.loads(l); print('  %3d replicas: %8.0f sweeps/s %7.1f attempts/ns  [%s]' % (d['replicas_on_rank'], d['rank_sweeps_per_s'], d['attempts_per_ns'], d['path']))
    else: print(l.strip()[:200])
PY
