#!/bin/bash
# round 2, GPU call 14 (2 GPUs): the default bench under torchrun at N = 2 after the barrier fix (timeout-bounded), digests against N = 1
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call14.log
: > $O
echo "== bench N=2" >> $O
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r02/call14_bench_n2.json 2> gpurun_out/r02/call14_bench_n2.err
tail -4 gpurun_out/r02/call14_bench_n2.err >> $O
echo "== bench N=1 (same box)" >> $O
( time timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu ) > gpurun_out/r02/call14_bench_n1.json 2> gpurun_out/r02/call14_bench_n1.err
tail -4 gpurun_out/r02/call14_bench_n1.err >> $O
echo "== --config c5 N=2 / --config c4 N=1" >> $O
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29822 bench.py --gpus 2 --config c5 --no-cpu ) > gpurun_out/r02/call14_bench_c5_n2.json 2>> $O
timeout 300 python bench.py --config c4 --no-cpu > gpurun_out/r02/call14_bench_c4_n1.json 2>> $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $O 2>&1
grep -v "OMP_NUM_THREADS\|\*\*\*\*" $O | cut -c1-260
