#!/bin/bash
# ticket-queue kernel, default policy: parity tests, then the per-rank PT rate at 32 / 64 replicas
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_queue.py -q -m gpu 2>&1 | tail -12 > gpurun_out/pytest_queue2.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_queue2.log
tail -4 gpurun_out/pytest_queue2.log
timeout 40 python scripts/bench_pt_rank.py --counts 32,64 --every 200 --rounds 5 > gpurun_out/queue_pt3.log 2>&1
cat gpurun_out/queue_pt3.log
