#!/bin/bash
# round 2, GPU call 10: pipe-balance probes of the thread-row update (shift / subtraction on the ALU instead of the FMA pipe),
# 3-D bands, Blume-Capel heat bath, PT batch on bit planes
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call10.log
: > $O
timeout 1700 python -m pytest tests/test_gpu_checkpoint.py tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_bits.py tests/test_gpu_full_oracle.py -x -q 2>&1 | tail -25 > gpurun_out/r02/call10_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02/call10_pytest.log
tail -12 gpurun_out/r02/call10_pytest.log
echo "== headline A/B: default, shr (shift on ALU), sub (plain subtraction), shrsub" >> $O
timeout 900 bash scripts/gpu_ab.sh default shr sub shrsub >> $O 2>&1
echo "== 3-D: bands off / default / 4 / 16" >> $O
MCX_BANDS=0 timeout 300 python scripts/bench_storage.py --sizes "" --d3 512,256,128 >> $O 2>&1
timeout 300 python scripts/bench_storage.py --sizes "" --d3 512,256,128 >> $O 2>&1
MCX_BANDS=4 timeout 300 python scripts/bench_storage.py --sizes "" --d3 256 >> $O 2>&1
MCX_BANDS=16 timeout 300 python scripts/bench_storage.py --sizes "" --d3 256 >> $O 2>&1
echo "== BC rates" >> $O
timeout 300 python scripts/bench_bc.py >> $O 2>&1
BC_RULE=heatbath timeout 300 python scripts/bench_bc.py >> $O 2>&1
cut -c1-260 $O
