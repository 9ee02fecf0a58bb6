#!/bin/bash
# round 2, GPU call 18: single-basic-block trip of k_ising2d (Philox head, deferred tie rows): parity of every path that
# runs k_ising2d, then an A/B against the library of the previous commit on this box
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call18.log
: > $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_oracle.py tests/test_gpu_full_size.py tests/test_gpu_slab.py tests/test_gpu_queue.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02/call18_pytest.log 2>&1
tail -6 gpurun_out/r02/call18_pytest.log
echo "== A/B" >> $O
bash scripts/gpu_ab.sh base default >> $O 2>&1
cat $O
