#!/bin/bash
# round 2, GPU call 9: general topologies
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_graphs.py -x -q 2>&1 | tail -30 > gpurun_out/r02/call9_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02/call9_pytest.log
tail -30 gpurun_out/r02/call9_pytest.log
