#!/bin/bash
# round 2, GPU call 23: launch shapes of k_ising2d at 8 CTAs/SM (bands x strip height)
mkdir -p gpurun_out/r02
timeout 600 python scripts/bench_bands.py > gpurun_out/r02/call23.log 2>&1
cat gpurun_out/r02/call23.log
