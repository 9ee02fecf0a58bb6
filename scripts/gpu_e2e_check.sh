#!/bin/bash
# Runs on the GPU box (via gpurun): split-upload test, golden trajectories, then the default bench line.
mkdir -p gpurun_out
timeout 200 python -m pytest "tests/test_gpu_parity.py::test_split_upload_equals_upload" "tests/test_gpu_parity.py::test_device_reproduces_committed_trajectories" -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_upload.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_upload.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit: $?" >> gpurun_out/bench_default.err
tail -8 gpurun_out/pytest_upload.log; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
