#!/bin/bash
# round 2, GPU call 16: final checks on one GPU -- full GPU suite, smoke, default bench + reference arm, ncu launch list of the bench
# command and one full capture of the headline kernel
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call16.log
: > $O
( time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02/call16_pytest.log 2>&1
tail -6 gpurun_out/r02/call16_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $O 2>&1
echo "== bench default" >> $O
( time timeout 1200 python bench.py ) > gpurun_out/r02/call16_bench.json 2> gpurun_out/r02/call16_bench.err
tail -4 gpurun_out/r02/call16_bench.err >> $O
echo "== ncu launch list" >> $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 1500 --csv --log-file gpurun_out/r02/launches.csv \
   python bench.py --steps 2 --warmup 3 --sweeps-per-step 10 --no-cpu --no-extras > gpurun_out/r02/call16_ncu_launches.log 2>&1
tail -2 gpurun_out/r02/call16_ncu_launches.log | cut -c1-200 >> $O
echo "== ncu full, headline kernel" >> $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ising2d -s 40 -c 2 -f -o gpurun_out/r02/ncu_ising2d_r02 \
   python bench.py --steps 1 --warmup 3 --sweeps-per-step 5 --no-pt --no-cpu --no-extras > gpurun_out/r02/call16_ncu_full.log 2>&1
tail -2 gpurun_out/r02/call16_ncu_full.log | cut -c1-200 >> $O
cut -c1-260 $O
