#!/bin/bash
# Runs on the GPU box (via gpurun): the windowed Wang-Landau parity tests, the golden-trajectory test, the C5
# windows timing and the smoke.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_windows.py "tests/test_gpu_parity.py::test_device_reproduces_committed_trajectories" -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_windows.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_windows.log
timeout 150 python scripts/bench_windows.py --L 256 --windows 8 --walkers 4 --sweeps 1 > gpurun_out/bench_windows.log 2>&1
echo "bench_windows exit: $?" >> gpurun_out/bench_windows.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -8 gpurun_out/pytest_windows.log; tail -3 gpurun_out/bench_windows.log; tail -2 gpurun_out/smoke.log
