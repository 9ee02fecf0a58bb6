#!/bin/bash
# round 2, GPU call 28: prefetch variants of the strip loop: L2 (default), L1, two rows further ahead
mkdir -p gpurun_out/r02
bash scripts/gpu_ab.sh default pfl1 pf2 pfl1a2 > gpurun_out/r02/call28.log 2>&1
cat gpurun_out/r02/call28.log
