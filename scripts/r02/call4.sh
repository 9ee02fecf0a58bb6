#!/bin/bash
# round 2, GPU call 4: static decomposition with boundary trips last, Philox before first use of the loaded rows
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call4.log
: > $O
timeout 1800 python -m pytest tests/test_gpu_queue.py tests/test_gpu_pt_persistent.py -x -q 2>&1 | tail -30 > gpurun_out/r02/call4_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02/call4_pytest.log
tail -5 gpurun_out/r02/call4_pytest.log
for st in 1 0; do
  echo "== every 200, MCX_PT_PERSIST=1 MCX_STATIC=$st" >> $O
  MCX_STATIC=$st MCX_PT_PERSIST=1 timeout 300 python scripts/bench_pt_rank.py --counts 256,128,64,32 >> $O 2>&1
  echo "== every 1, MCX_PT_PERSIST=1 MCX_STATIC=$st" >> $O
  MCX_STATIC=$st MCX_PT_PERSIST=1 timeout 300 python scripts/bench_pt_rank.py --counts 256,128,64,32 --every 1 --rounds 300 >> $O 2>&1
done
for rows in 10 12 16 22; do
  echo "== every 200, static rows=$rows" >> $O
  MCX_QUEUE_ROWS=$rows MCX_PT_PERSIST=1 timeout 300 python scripts/bench_pt_rank.py --counts 64,32 >> $O 2>&1
  echo "== every 1, static rows=$rows" >> $O
  MCX_QUEUE_ROWS=$rows MCX_PT_PERSIST=1 timeout 300 python scripts/bench_pt_rank.py --counts 64,32 --every 1 --rounds 300 >> $O 2>&1
done
echo "== headline: default vs philox-first" >> $O
bash scripts/gpu_ab.sh default pf >> $O 2>&1
echo "== single lattice static (MCX_QUEUE=1)" >> $O
MCX_QUEUE=1 timeout 300 python bench.py --no-cpu --no-pt --steps 3 --warmup 3 >> $O 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ising2d_static -s 1 -c 1 -f -o gpurun_out/r02/ncu_static2 \
   env MCX_PT_PERSIST=1 python scripts/bench_pt_rank.py --counts 32 --every 20 --rounds 2 > gpurun_out/r02/ncu_static2.log 2>&1
grep -c replicas_on_rank $O
