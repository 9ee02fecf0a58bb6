#!/bin/bash
# round 2, GPU call 7: flat-histogram windows (parity + rates), the new bench legs
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call7.log
: > $O
timeout 1500 python -m pytest tests/test_gpu_flat.py tests/test_gpu_windows.py -x -q 2>&1 | tail -30 > gpurun_out/r02/call7_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02/call7_pytest.log
tail -15 gpurun_out/r02/call7_pytest.log
echo "== flat rates, window on" >> $O
timeout 600 python scripts/bench_flat.py c4 c5 muca2d >> $O 2>&1
echo "== flat rates, MCX_FLAT_WINDOW=0" >> $O
MCX_FLAT_WINDOW=0 timeout 600 python scripts/bench_flat.py c4 c5 muca2d >> $O 2>&1
echo "== WL scan (adaptive)" >> $O
timeout 600 python scripts/bench_flat.py wlscan 2>&1 | grep adaptive >> $O
echo "== bench default" >> $O
( time timeout 1200 python bench.py ) > gpurun_out/r02/call7_bench.json 2> gpurun_out/r02/call7_bench.err
tail -5 gpurun_out/r02/call7_bench.err >> $O
echo "== bench --config c5" >> $O
( time timeout 900 python bench.py --config c5 ) > gpurun_out/r02/call7_bench_c5.json 2> gpurun_out/r02/call7_bench_c5.err
tail -5 gpurun_out/r02/call7_bench_c5.err >> $O
cut -c1-300 $O
