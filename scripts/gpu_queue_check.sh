#!/bin/bash
# Runs on the GPU box (via gpurun): ticket-queue series kernel -- parity tests, then the per-rank PT rate and the
# headline lattice with and without MCX_QUEUE=1.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_queue.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/pytest_queue.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_queue.log
tail -6 gpurun_out/pytest_queue.log
for q in "" 1; do
  echo "== MCX_QUEUE=$q" >> gpurun_out/queue_pt.log
  MCX_QUEUE=$q timeout 120 python scripts/bench_pt_rank.py --counts 256,128,64,32 >> gpurun_out/queue_pt.log 2>&1
  MCX_QUEUE=$q timeout 120 python bench.py --no-pt --no-cpu --steps 3 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    print(json.dumps({'value': d['value'], 'kernel_attempts_per_ns': d['roofline']['kernel_attempts_per_ns'], 'launches': d['gpu_launches'], 'e2e': d['e2e']['value']}))
" >> gpurun_out/queue_pt.log
done
cat gpurun_out/queue_pt.log
