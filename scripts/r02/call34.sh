#!/bin/bash
# round 2, GPU call 34: parallel-tempering rounds replayed from a CUDA graph (device clock): parity against the rounds queued launch
# by launch, per-rank rates with and without it; tracked sums with the dp4a accounting
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call34.log
: > $O
( timeout 900 python -m pytest tests/test_gpu_pt_persistent.py tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r02/call34_pytest.log 2>&1
tail -3 gpurun_out/r02/call34_pytest.log
for g in 0 1; do
  echo "== MCX_PT_GRAPH=$g every 1" >> $O
  MCX_PT_GRAPH=$g timeout 300 python scripts/bench_pt_rank.py --counts 256,64,32 --every 1 --rounds 600 2>&1 | python -c "
import sys, re
for l in sys.stdin:
    if l.startswith('{'):
        g = lambda k: re.search(r'\"%s\": ([^,}]+)' % k, l).group(1)
        print('  %3s replicas: %9.0f sweeps/s %7.1f attempts/ns host %s us/sweep [%s]' % (g('replicas_on_rank'), float(g('rank_sweeps_per_s')), float(g('attempts_per_ns')), g('host_enqueue_us_per_sweep'), g('path')))
    else: print(l.strip()[:200])" >> $O
done
echo "== MCX_PT_PERSIST=0 (graph), 64 and 32 replicas, every 1 / 2" >> $O
for ev in 1 2; do MCX_PT_PERSIST=0 timeout 300 python scripts/bench_pt_rank.py --counts 64,32 --every $ev --rounds 600 2>&1 | python -c "
import sys, re
for l in sys.stdin:
    if l.startswith('{'):
        g = lambda k: re.search(r'\"%s\": ([^,}]+)' % k, l).group(1)
        print('  every $ev %3s replicas: %9.0f sweeps/s %7.1f attempts/ns host %s us/sweep [%s]' % (g('replicas_on_rank'), float(g('rank_sweeps_per_s')), float(g('attempts_per_ns')), g('host_enqueue_us_per_sweep'), g('path')))
    else: print(l.strip()[:200])" >> $O
done
echo "== tracked sums (bench --track 1)" >> $O
for v in base default; do
  if [ "$v" = default ]; then unset MCX_B200_LIB; else export MCX_B200_LIB=$PWD/montecarlox.jl_b200/lib/libmcx_b200_$v.so; fi
  timeout 300 python bench.py --track 1 --no-cpu --no-pt --no-extras --steps 3 --warmup 3 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('tracked LIB=$v value=%.1f kernel=%.1f' % (d['value'], d['roofline']['kernel_attempts_per_ns']))" >> $O
done
cat $O
