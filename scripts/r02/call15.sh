#!/bin/bash
# round 2, GPU call 15 (8 GPUs): the scaling run's own commands once, bounded by timeouts
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call15.log
: > $O
nvidia-smi -L >> $O 2>&1
nvidia-smi topo -m >> $O 2>&1
echo "== bench N=8" >> $O
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r02/call15_bench_n8.json 2> gpurun_out/r02/call15_bench_n8.err
tail -4 gpurun_out/r02/call15_bench_n8.err >> $O
echo "== windows with neighbour exchanges over NCCL, N=8" >> $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29832 scripts/nccl_windows_check.py 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -5 >> $O
echo "== bench --config c3 N=8" >> $O
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29833 bench.py --gpus 8 --config c3 --no-cpu ) > gpurun_out/r02/call15_bench_c3_n8.json 2>> $O
echo "== bench N=4" >> $O
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29834 bench.py --gpus 4 --steps 3 --warmup 3 ) > gpurun_out/r02/call15_bench_n4.json 2> gpurun_out/r02/call15_bench_n4.err
tail -4 gpurun_out/r02/call15_bench_n4.err >> $O
grep -v "OMP_NUM_THREADS\|\*\*\*\*" $O | cut -c1-260
