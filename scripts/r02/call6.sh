#!/bin/bash
# round 2, GPU call 6: one-bit storage -- parity tests, then int8 vs bit rates
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call6.log
: > $O
timeout 1500 python -m pytest tests/test_gpu_bits.py -x -q 2>&1 | tail -30 > gpurun_out/r02/call6_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02/call6_pytest.log
tail -15 gpurun_out/r02/call6_pytest.log
echo "== default policy" >> $O
timeout 600 python scripts/bench_storage.py >> $O 2>&1
echo "== MCX_BANDS=0" >> $O
MCX_BANDS=0 timeout 300 python scripts/bench_storage.py --sizes 16384,8192 --d3 "" >> $O 2>&1
echo "== PT batch 32 / 256 x 1024^2" >> $O
timeout 300 python scripts/bench_storage.py --sizes 1024 --d3 "" --chains 32 >> $O 2>&1
timeout 300 python scripts/bench_storage.py --sizes 1024 --d3 "" --chains 256 --sweeps 20 >> $O 2>&1
cat $O | cut -c1-200
