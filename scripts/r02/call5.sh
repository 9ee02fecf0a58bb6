#!/bin/bash
# round 2, GPU call 5: r01 ticket queue restored for plain series; persistent PT rounds on the static warp-strip kernel
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call5.log
: > $O
timeout 1800 python -m pytest tests/test_gpu_queue.py tests/test_gpu_pt_persistent.py -x -q 2>&1 | tail -30 > gpurun_out/r02/call5_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r02/call5_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_checkpoint.py -x -q 2>&1 | tail -8 >> gpurun_out/r02/call5_pytest.log
tail -12 gpurun_out/r02/call5_pytest.log
echo "== every 200 default policy" >> $O
timeout 300 python scripts/bench_pt_rank.py --counts 256,128,64,32 >> $O 2>&1
echo "== every 1 default policy" >> $O
timeout 300 python scripts/bench_pt_rank.py --counts 256,128,64,32 --every 1 --rounds 300 >> $O 2>&1
echo "== every 1 host-queued" >> $O
MCX_PT_PERSIST=0 timeout 300 python scripts/bench_pt_rank.py --counts 256,32 --every 1 --rounds 300 >> $O 2>&1
for rows in 12 16 22; do
  echo "== every 1, rows=$rows" >> $O
  MCX_QUEUE_ROWS=$rows timeout 300 python scripts/bench_pt_rank.py --counts 64,32 --every 1 --rounds 300 >> $O 2>&1
done
echo "== every 2 / 4 / 8 default" >> $O
for e in 2 4 8; do timeout 300 python scripts/bench_pt_rank.py --counts 256,32 --every $e --rounds 100 >> $O 2>&1; done
echo "== every 4 host-queued" >> $O
MCX_PT_PERSIST=0 timeout 300 python scripts/bench_pt_rank.py --counts 256,32 --every 4 --rounds 100 >> $O 2>&1
echo "== every 200 persistent forced" >> $O
MCX_PT_PERSIST=1 timeout 300 python scripts/bench_pt_rank.py --counts 256,32 >> $O 2>&1
grep -c replicas_on_rank $O
