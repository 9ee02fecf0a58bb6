#!/bin/bash
# round 2, GPU call 13 (2 GPUs): parity on distinct devices, the bench under torchrun at N = 2 (configs c4 / c5 with their collectives),
# and the N = 1 line on the same box for the parity digests
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call13.log
: > $O
nvidia-smi -L >> $O 2>&1
( time timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slab.py tests/test_gpu_windows.py -x -q 2>&1 | tail -15 ) > gpurun_out/r02/call13_pytest.log 2>&1
tail -8 gpurun_out/r02/call13_pytest.log
echo "== bench N=2" >> $O
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r02/call13_bench_n2.json 2> gpurun_out/r02/call13_bench_n2.err
tail -4 gpurun_out/r02/call13_bench_n2.err >> $O
echo "== bench N=1" >> $O
( time timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu ) > gpurun_out/r02/call13_bench_n1.json 2> gpurun_out/r02/call13_bench_n1.err
tail -4 gpurun_out/r02/call13_bench_n1.err >> $O
echo "== bench --config c4 N=2, --config c3 N=2" >> $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 2 --config c4 --no-cpu > gpurun_out/r02/call13_bench_c4_n2.json 2>> $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29813 bench.py --gpus 2 --config c3 --no-cpu > gpurun_out/r02/call13_bench_c3_n2.json 2>> $O
cut -c1-260 $O
