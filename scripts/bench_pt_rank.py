"""Per-rank parallel-tempering rate on ONE GPU for the replica counts a rank holds at 1/2/4/8 GPUs.

    python scripts/bench_pt_rank.py [--L 1024] [--counts 256,128,64,32] [--every 200]

Strong scaling of BASELINE.json configs[2] leaves a rank with 256/N replicas; this times exactly that
share (sweeps + label exchange, no collective) so that small-grid effects can be tuned without an N-GPU box.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=1024)
    ap.add_argument("--counts", default="256,128,64,32")
    ap.add_argument("--every", type=int, default=200)
    ap.add_argument("--rounds", type=int, default=5)
    args = ap.parse_args()
    import torch
    import mcx_b200 as m
    stream = torch.cuda.Stream()
    ctx = m.Context(0, stream=stream.cuda_stream)
    with torch.cuda.stream(stream):
        for count in [int(c) for c in args.counts.split(",")]:
            betas = m.set_betas(count, 1 / 3.0, 1 / 1.5, "uniform")
            pt = m.ParallelTempering(betas, seed=42, backend=m.GPUBackend())
            reps = m.Ising([args.L, args.L], nchains=count, ctx=ctx)
            pt.attach(reps)
            reps.init_("random", rng=m.PhiloxRNG(42, 0))

            def pt_round():
                m.sweep_(reps, pt, args.every)
                m.update_(pt)

            for _ in range(3):
                pt_round()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            import time
            e0.record(stream)
            h0 = time.perf_counter()
            if os.environ.get("PT_LOOP", "run") == "run":
                pt.run_(reps, args.rounds, args.every)
            else:
                for _ in range(args.rounds):
                    pt_round()
            host_us = (time.perf_counter() - h0) * 1e6 / (args.rounds * args.every)    # host time to ENQUEUE a sweep
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            sweeps = args.rounds * args.every
            import ctypes as C
            path, strip = C.c_int32(-1), C.c_int32(0)
            m.lib().mcx_pt_run_info(pt._pt, C.byref(path), C.byref(strip))
            print(json.dumps({"path": "persistent launch, %d-row strips" % strip.value if path.value == 1 else "graph replay" if path.value == 2 else "host-queued rounds","replicas_on_rank": count, "L": args.L, "every": args.every,
                              "rank_sweeps_per_s": sweeps / (ms * 1e-3), "us_per_sweep": ms * 1e3 / sweeps, "host_enqueue_us_per_sweep": round(host_us, 2),
                              "attempts_per_ns": sweeps * count * args.L * args.L / (ms * 1e6),
                              "rows_per_strip": os.environ.get("MCX_ROWS_PER_STRIP", "auto")}), flush=True)
            pt.close() if hasattr(pt, "close") else None
            del reps, pt


if __name__ == "__main__":
    main()
