#!/bin/bash
# round 2, GPU call 11: full GPU suite after the pipe-balance change; A/B default vs IMAD.HI shift vs unrolled trips;
# ticket-queue series for single mid-size lattices; 3-D and Blume-Capel rates
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call11.log
: > $O
( time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02/call11_pytest.log 2>&1
tail -8 gpurun_out/r02/call11_pytest.log
echo "== headline A/B: default (ALU shift), shrfma (round-1 IMAD.HI shift), unroll2" >> $O
timeout 900 bash scripts/gpu_ab.sh default shrfma unroll2 >> $O 2>&1
echo "== single mid-size lattices: default policy" >> $O
timeout 300 python scripts/bench_storage.py --sizes 8192,4096,2048,1024 --d3 "" 2>&1 | grep '"int8", "track": 0' >> $O
for rows in 2 4 8 16; do
  echo "== MCX_QUEUE=1 MCX_QUEUE_ROWS=$rows" >> $O
  MCX_QUEUE=1 MCX_QUEUE_ROWS=$rows timeout 300 python scripts/bench_storage.py --sizes 8192,4096,2048,1024 --d3 "" 2>&1 | grep '"int8", "track": 0' >> $O
done
echo "== 3-D default" >> $O
timeout 300 python scripts/bench_storage.py --sizes "" --d3 512,256 2>&1 | grep '"track": 0' >> $O
echo "== BC rates" >> $O
timeout 300 python scripts/bench_bc.py >> $O 2>&1
BC_RULE=heatbath timeout 300 python scripts/bench_bc.py >> $O 2>&1
cut -c1-260 $O
