#!/bin/bash
# round 2, GPU call 8: default bench with the new legs; new parity tests; ncu of the flat-histogram kernel on C4 and of the bit half-sweep
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call8.log
: > $O
timeout 1500 python -m pytest tests/test_gpu_graphs.py tests/test_gpu_checkpoint.py tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_full_oracle.py -x -q 2>&1 | tail -40 > gpurun_out/r02/call8_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02/call8_pytest.log
tail -30 gpurun_out/r02/call8_pytest.log
echo "== bench default" >> $O
( time timeout 1200 python bench.py ) > gpurun_out/r02/call8_bench.json 2> gpurun_out/r02/call8_bench.err
tail -5 gpurun_out/r02/call8_bench.err >> $O
echo "== BC rates" >> $O
timeout 300 python scripts/bench_bc.py >> $O 2>&1
BC_RULE=heatbath timeout 300 python scripts/bench_bc.py >> $O 2>&1
echo "== ncu k_flat_warp (C4)" >> $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_flat_warp -s 1 -c 1 -f -o gpurun_out/r02/ncu_flat_c4 \
   python scripts/bench_flat.py c4 > gpurun_out/r02/call8_ncu_flat.log 2>&1
tail -3 gpurun_out/r02/call8_ncu_flat.log >> $O
echo "== ncu k_ising2d_bits" >> $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ising2d_bits -s 20 -c 2 -f -o gpurun_out/r02/ncu_bits \
   python scripts/bench_storage.py --sizes 16384 --d3 "" --sweeps 10 > gpurun_out/r02/call8_ncu_bits.log 2>&1
tail -3 gpurun_out/r02/call8_ncu_bits.log >> $O
ls -la gpurun_out/r02 >> $O
cut -c1-300 $O
