"""Sweep-series rate of small 2-D Ising batches: streaming half-sweep launches vs the shared-memory
resident series kernel (MCX_RESIDENT=0 / 1).  python scripts/bench_small.py [--sweeps 200]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweeps", type=int, default=200)
    ap.add_argument("--shapes", default="64x1,64x64,64x4096,128x1,128x256,256x1,256x16,256x256,512x1,512x8,512x64,1024x1,1024x8")
    args = ap.parse_args()
    import torch
    import mcx_b200 as m
    stream = torch.cuda.Stream()
    ctx = m.Context(0, stream=stream.cuda_stream)
    for shape in args.shapes.split(","):
        L, n = [int(v) for v in shape.split("x")]
        row = {"L": L, "chains": n}
        for mode in ("0", "1"):
            os.environ["MCX_RESIDENT"] = mode
            sys_ = m.Ising([L, L], nchains=n, ctx=ctx)
            sys_.set_tracking(False)
            rng = m.PhiloxRNG(3)
            alg = m.Metropolis(rng, beta=0.44)
            m.init_(sys_, "random", rng=rng)
            for _ in range(2):
                m.sweep_(sys_, alg, args.sweeps)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(3):
                m.sweep_(sys_, alg, args.sweeps)
            e1.record(stream)
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (3 * args.sweeps)
            row["us_per_sweep_" + ("resident" if mode == "1" else "streaming")] = round(us, 3)
            row["attempts_per_ns_" + ("resident" if mode == "1" else "streaming")] = round(n * L * L / (us * 1e3), 2)
            del sys_
        row["speedup"] = round(row["us_per_sweep_streaming"] / row["us_per_sweep_resident"], 2)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
