#!/bin/bash
# ticket-queue kernel: exchange after every sweep (host-launch-bound regime) and strip heights at 32 replicas
mkdir -p gpurun_out
run() { echo "== $*" >> gpurun_out/queue_pt2.log; env "$@" PT_LOOP=run timeout 100 python scripts/bench_pt_rank.py --counts $COUNTS --every $EVERY --rounds $ROUNDS >> gpurun_out/queue_pt2.log 2>&1; }
COUNTS=32,64 EVERY=1 ROUNDS=600
run MCX_QUEUE=0
run MCX_QUEUE=1
COUNTS=32 EVERY=200 ROUNDS=5
run MCX_QUEUE=1 MCX_QUEUE_ROWS=16
run MCX_QUEUE=1 MCX_QUEUE_ROWS=8
cat gpurun_out/queue_pt2.log
