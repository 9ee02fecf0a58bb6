#!/bin/bash
# round 2, GPU call 27: one L2 prefetch instruction per trip (default) against the same kernel without it (nopf); parity
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call27.log
: > $O
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_oracle.py tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r02/call27_pytest.log 2>&1
tail -2 gpurun_out/r02/call27_pytest.log
bash scripts/gpu_ab.sh nopf default >> $O 2>&1
cat $O
