#!/bin/bash
# round 2, GPU call 19: k_ising2d single-basic-block trip at 6 / 7 / 8 CTAs per SM (80 / 72 / 64 registers); parity first
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call19.log
: > $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_oracle.py tests/test_gpu_full_size.py tests/test_gpu_slab.py tests/test_gpu_queue.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02/call19_pytest.log 2>&1
tail -6 gpurun_out/r02/call19_pytest.log
echo "== A/B" >> $O
bash scripts/gpu_ab.sh base default mb7 mb8 >> $O 2>&1
cat $O
