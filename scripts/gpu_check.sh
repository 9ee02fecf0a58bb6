#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, smoke, a short bench.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --sweeps-per-step 20 --cpu-seconds 5 > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log
if [ "$1" == "prof" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ising2d -s 12 -c 2 -f -o gpurun_out/prof_ising2d \
     python bench.py --steps 1 --warmup 1 --sweeps-per-step 5 --no-pt --no-cpu > gpurun_out/ncu_full.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 200 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 2 --warmup 1 --sweeps-per-step 10 --no-cpu > gpurun_out/ncu_launches.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
