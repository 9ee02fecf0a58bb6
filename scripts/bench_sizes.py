"""Sweep rate of single 2-D Ising lattices of several sizes, plain launches vs row bands (MCX_BANDS=0 / default)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcx_b200 as m

stream = torch.cuda.Stream()
ctx = m.Context(0, stream=stream.cuda_stream)
for L in [int(v) for v in os.environ.get("SIZES", "2048,4096,8192,16384,32768").split(",")]:
    row = {"L": L}
    for mode, env in (("plain", {"MCX_BANDS": "0"}), ("bands", {} if "BANDS" not in os.environ else {"MCX_BANDS": os.environ["BANDS"]})):
        os.environ.pop("MCX_BANDS", None); os.environ.update(env)
        s = m.Ising([L, L], ctx=ctx); s.set_tracking(False)
        rng = m.PhiloxRNG(3); alg = m.Metropolis(rng, beta=0.44)
        m.init_(s, "random", rng=rng)
        n = max(4, int(2e10 / (L * L)) // 4)
        m.sweep_(s, alg, n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        l0 = ctx.launch_count()
        from mcx_b200._lib import check, lib
        check(lib().mcx_sweep(s.h_lat, n))
        e1.record(stream); torch.cuda.synchronize()
        row[mode] = round(n * L * L / (e0.elapsed_time(e1) * 1e6), 1)
        row[mode + "_launches_per_sweep"] = (ctx.launch_count() - l0) / n
        del s
    print(json.dumps(row), flush=True)
