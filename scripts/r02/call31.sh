#!/bin/bash
# round 2, GPU call 31: 3-D Ising kernel on the one-block-per-trip loop: parity, then rates at 5 / 6 / 8 CTAs per SM against
# the previous round's library
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call31.log
: > $O
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bits.py -m gpu -x -q -k "3d or 3D or bands or trajector" 2>&1 | tail -4 ) > gpurun_out/r02/call31_pytest.log 2>&1
tail -2 gpurun_out/r02/call31_pytest.log
( timeout 300 python -m pytest tests/test_gpu_golden.py tests/test_gpu_statistics.py -m gpu -x -q 2>&1 | tail -2 ) >> gpurun_out/r02/call31_pytest.log 2>&1
tail -1 gpurun_out/r02/call31_pytest.log
for v in base default 3d6 3d8; do
  if [ "$v" = default ]; then unset MCX_B200_LIB; else export MCX_B200_LIB=$PWD/montecarlox.jl_b200/lib/libmcx_b200_$v.so; fi
  echo "== LIB=$v" >> $O
  timeout 300 python scripts/bench_3d.py 2>&1 | cut -c1-200 >> $O
done
cat $O
