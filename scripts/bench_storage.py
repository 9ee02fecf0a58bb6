"""attempts/ns of the checkerboard half-sweeps on int8 and on one-bit planes (MCX_STORAGE_BIT), 2-D and 3-D.

    python scripts/bench_storage.py [--sizes 16384,8192,4096] [--sweeps 50] [--d3 512,256]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BETA_C = 0.440686793509772


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="16384,8192,4096,2048")
    ap.add_argument("--d3", default="512,256")
    ap.add_argument("--sweeps", type=int, default=50)
    ap.add_argument("--rule", type=int, default=0)
    ap.add_argument("--chains", type=int, default=1)
    args = ap.parse_args()
    import torch
    import mcx_b200 as m
    from mcx_b200._lib import check, lib
    stream = torch.cuda.Stream()
    ctx = m.Context(0, stream=stream.cuda_stream)
    shapes = [[int(L), int(L)] for L in args.sizes.split(",") if L] + [[int(L)] * 3 for L in args.d3.split(",") if L]
    with torch.cuda.stream(stream):
        for dims in shapes:
            for storage in ("int8", "bit"):
                for track in (0, 1):
                    s = m.Ising(dims, nchains=args.chains, ctx=ctx, storage=storage)
                    s.set_tracking(bool(track))
                    rng = m.PhiloxRNG(42, 0)
                    alg = (m.Metropolis, m.Glauber, m.HeatBath)[args.rule](rng, beta=BETA_C if len(dims) == 2 else 0.2216544)
                    s._bind_alg(alg)
                    s.init_("random", rng=rng)
                    check(lib().mcx_sweep(s.h_lat, 5))
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    l0 = ctx.launch_count()
                    e0.record(stream)
                    check(lib().mcx_sweep(s.h_lat, args.sweeps))
                    e1.record(stream)
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1)
                    n = args.chains * int(torch.tensor(dims).prod())
                    print(json.dumps({"dims": dims, "chains": args.chains, "storage": storage, "track": track, "rule": args.rule,
                                      "attempts_per_ns": round(args.sweeps * n / (ms * 1e6), 1),
                                      "us_per_half_sweep": round(ms * 1e3 / (2 * args.sweeps), 2),
                                      "launches_per_half_sweep": (ctx.launch_count() - l0) / (2 * args.sweeps),
                                      "env": {k: v for k, v in os.environ.items() if k.startswith("MCX_")}}), flush=True)
                    del s


if __name__ == "__main__":
    main()
