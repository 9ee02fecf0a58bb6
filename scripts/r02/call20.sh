#!/bin/bash
# round 2, GPU call 20: k_ising2d at 8 / 9 / 10 / 12 CTAs per SM (64 / 56 / 48 / 40 registers)
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call20.log
: > $O
bash scripts/gpu_ab.sh default mb9 mb10 mb12 >> $O 2>&1
cat $O
