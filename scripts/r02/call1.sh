#!/bin/bash
# round 2, GPU call 1: queue kernel at 5 vs 6 CTAs/SM on the replica shares of 1/2/4/8 GPUs, one ncu capture of it,
# and the tracked-sums rate of the headline lattice
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call1.log
: > $O
Q6=$PWD/montecarlox.jl_b200/lib/libmcx_b200_q6.so
for rep in 1 2; do
  echo "== default policy (shipped lib)" >> $O
  python scripts/bench_pt_rank.py --counts 256,128,64,32 >> $O 2>&1
  for rows in 8 16; do
    echo "== q5 MCX_QUEUE=1 rows=$rows" >> $O
    MCX_QUEUE=1 MCX_QUEUE_ROWS=$rows python scripts/bench_pt_rank.py --counts 256,128,64,32 >> $O 2>&1
    echo "== q6 MCX_QUEUE=1 rows=$rows" >> $O
    MCX_B200_LIB=$Q6 MCX_QUEUE=1 MCX_QUEUE_ROWS=$rows python scripts/bench_pt_rank.py --counts 256,128,64,32 >> $O 2>&1
  done
  echo "== q6 MCX_QUEUE=1 rows=4 (32, 64)" >> $O
  MCX_B200_LIB=$Q6 MCX_QUEUE=1 MCX_QUEUE_ROWS=4 python scripts/bench_pt_rank.py --counts 64,32 >> $O 2>&1
done
echo "== every sweep, default / q6 queue" >> $O
python scripts/bench_pt_rank.py --counts 256,32 --every 1 --rounds 300 >> $O 2>&1
echo "== tracked headline" >> $O
python bench.py --no-cpu --no-pt --steps 3 --warmup 3 --track 1 >> $O 2>&1
python bench.py --no-cpu --no-pt --steps 3 --warmup 3 --track 0 >> $O 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ising2d_queue -s 3 -c 1 -f -o gpurun_out/r02/ncu_queue_q5 \
   python scripts/bench_pt_rank.py --counts 32 --every 20 --rounds 2 > gpurun_out/r02/ncu_queue_q5.log 2>&1
MCX_B200_LIB=$Q6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ising2d_queue -s 3 -c 1 -f -o gpurun_out/r02/ncu_queue_q6 \
   python scripts/bench_pt_rank.py --counts 32 --every 20 --rounds 2 > gpurun_out/r02/ncu_queue_q6.log 2>&1
tail -40 $O
