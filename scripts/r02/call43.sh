#!/bin/bash
# round 2, GPU call 43: compute-sanitizer memcheck on the reworked strip loop (deferred rows, prefetch past the strip, bit planes,
# series kernels) and on the graph replays
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call43.log
: > $O
echo "== memcheck: streaming kernel, bands, groups, sweep graph, bit planes" >> $O
( time timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bits.py -m gpu -x -q -k "ising2d_bit_exact or fast_kernel or bands_and_groups or graph_replayed or bits_2d or bit_planes" 2>&1 | tail -6 ) >> $O 2>&1
echo "== memcheck: ticket queue, persistent rounds, PT graph" >> $O
( time timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_queue.py tests/test_gpu_pt_persistent.py -m gpu -x -q -k "queue or graph_replayed or (persistent_rounds_equal and 1024)" 2>&1 | tail -6 ) >> $O 2>&1
cat $O
