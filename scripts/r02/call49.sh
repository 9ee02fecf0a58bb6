#!/bin/bash
# round 2, GPU call 49: ncu --set full of one whole-lattice launch of the final k_ising2d (two trips per iteration, L2 prefetch)
mkdir -p gpurun_out/r02
MCX_BANDS=0 timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_ising2d -s 8 -c 2 -f -o gpurun_out/r02/ncu_ising2d_v11_plain \
  python bench.py --steps 1 --warmup 1 --sweeps-per-step 5 --no-pt --no-cpu --no-extras > gpurun_out/r02/call49_ncu.log 2>&1
tail -2 gpurun_out/r02/call49_ncu.log | cut -c1-200
ls -la gpurun_out/r02/ | grep v11
