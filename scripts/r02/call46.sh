#!/bin/bash
# round 2, GPU call 46: the full GPU suite and smoke on the final tree
mkdir -p gpurun_out/r02
( time timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r02/call46_pytest.log 2>&1
tail -5 gpurun_out/r02/call46_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
