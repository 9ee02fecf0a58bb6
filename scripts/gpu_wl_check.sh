#!/bin/bash
# Runs on the GPU box (via gpurun): flat-histogram + windows parity tests, then the Wang-Landau width scan
# (scripts/bench_flat.py wlscan).  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_flat.py tests/test_gpu_windows.py -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_flat.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_flat.log
timeout 200 python scripts/bench_flat.py wlscan > gpurun_out/wlscan.log 2>&1
echo "wlscan exit: $?" >> gpurun_out/wlscan.log
tail -8 gpurun_out/pytest_flat.log; cat gpurun_out/wlscan.log
