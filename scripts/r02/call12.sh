#!/bin/bash
# round 2, GPU call 12: full GPU suite, smoke, default bench; mid-size lattices under the new default policy; 3-D Blume-Capel
mkdir -p gpurun_out/r02
O=gpurun_out/r02/call12.log
: > $O
( time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02/call12_pytest.log 2>&1
tail -8 gpurun_out/r02/call12_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $O 2>&1
echo "== bench default" >> $O
( time timeout 1200 python bench.py ) > gpurun_out/r02/call12_bench.json 2> gpurun_out/r02/call12_bench.err
tail -5 gpurun_out/r02/call12_bench.err >> $O
echo "== bench --impl reference" >> $O
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02/call12_bench_ref.json 2>> $O
echo "== mid-size lattices, default policy" >> $O
timeout 300 python scripts/bench_storage.py --sizes 8192,4096,2048,1024 --d3 "" 2>&1 | grep '"int8"' >> $O
echo "== 3-D Blume-Capel" >> $O
timeout 300 python scripts/bench_bc3d.py >> $O 2>&1
cut -c1-260 $O
