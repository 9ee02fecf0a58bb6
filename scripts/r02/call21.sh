#!/bin/bash
# round 2, GPU call 21: ncu --set full of the k_ising2d row-band launch (8 CTAs/SM build)
mkdir -p gpurun_out/r02
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ising2d -s 40 -c 2 -f -o gpurun_out/r02/ncu_ising2d_v8 \
   python bench.py --steps 1 --warmup 1 --sweeps-per-step 5 --no-pt --no-cpu --no-extras > gpurun_out/r02/call21_ncu.log 2>&1
tail -3 gpurun_out/r02/call21_ncu.log
ls -la gpurun_out/r02/
