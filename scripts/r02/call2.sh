#!/bin/bash
# round 2, GPU call 2: warp-granular ticket queue + persistent parallel-tempering rounds: parity, then rates
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call2.log
: > $O
timeout 1500 python -m pytest tests/test_gpu_queue.py tests/test_gpu_pt_persistent.py -x -q 2>&1 | tail -30 > gpurun_out/r02/call2_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02/call2_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -x -q -k "pt or tempering or peer" 2>&1 | tail -15 >> gpurun_out/r02/call2_pytest.log
tail -5 gpurun_out/r02/call2_pytest.log
for rows in auto 4 6 8 10 12 16; do
  echo "== every 200, MCX_PT_PERSIST=1 rows=$rows" >> $O
  if [ $rows = auto ]; then unset MCX_QUEUE_ROWS; else export MCX_QUEUE_ROWS=$rows; fi
  MCX_PT_PERSIST=1 timeout 300 python scripts/bench_pt_rank.py --counts 256,64,32 >> $O 2>&1
done
unset MCX_QUEUE_ROWS
echo "== every 200, default policy" >> $O
timeout 300 python scripts/bench_pt_rank.py --counts 256,128,64,32 >> $O 2>&1
echo "== every 200, MCX_PT_PERSIST=0 (host-queued rounds)" >> $O
MCX_PT_PERSIST=0 timeout 300 python scripts/bench_pt_rank.py --counts 256,128,64,32 >> $O 2>&1
for rows in auto 4 6 8 10 16; do
  echo "== every 1, MCX_PT_PERSIST=1 rows=$rows" >> $O
  if [ $rows = auto ]; then unset MCX_QUEUE_ROWS; else export MCX_QUEUE_ROWS=$rows; fi
  MCX_PT_PERSIST=1 timeout 300 python scripts/bench_pt_rank.py --counts 256,64,32 --every 1 --rounds 300 >> $O 2>&1
done
unset MCX_QUEUE_ROWS
echo "== every 1, MCX_PT_PERSIST=0" >> $O
MCX_PT_PERSIST=0 timeout 300 python scripts/bench_pt_rank.py --counts 256,32 --every 1 --rounds 300 >> $O 2>&1
echo "== single lattice through the queue (MCX_QUEUE=1) vs bands" >> $O
MCX_QUEUE=1 timeout 300 python bench.py --no-cpu --no-pt --steps 3 --warmup 3 >> $O 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ising2d_queue -s 1 -c 1 -f -o gpurun_out/r02/ncu_queue_warp \
   env MCX_PT_PERSIST=1 python scripts/bench_pt_rank.py --counts 32 --every 20 --rounds 2 > gpurun_out/r02/ncu_queue_warp.log 2>&1
grep -c replicas_on_rank $O
