"""Headline lattice (2-D Ising L = 16384, Metropolis at beta_c, int8) under different launch shapes of k_ising2d:
row bands x strip height (MCX_BANDS, MCX_BAND_ROWS), and plain launches (MCX_BANDS=0, MCX_ROWS_PER_STRIP)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcx_b200 as m
from mcx_b200._lib import check, lib

L = int(os.environ.get("L", "16384"))
stream = torch.cuda.Stream()
ctx = m.Context(0, stream=stream.cuda_stream)
shapes = [{}, {"MCX_BAND_ROWS": "32"}, {"MCX_BAND_ROWS": "8"}, {"MCX_BANDS": "16"}, {"MCX_BANDS": "16", "MCX_BAND_ROWS": "32"},
          {"MCX_BANDS": "4", "MCX_BAND_ROWS": "32"}, {"MCX_BANDS": "0"}, {"MCX_BANDS": "0", "MCX_ROWS_PER_STRIP": "32"}, {}]
keys = ("MCX_BANDS", "MCX_BAND_ROWS", "MCX_ROWS_PER_STRIP")
for env in shapes:
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    s = m.Ising([L, L], ctx=ctx); s.set_tracking(False)
    rng = m.PhiloxRNG(3); alg = m.Metropolis(rng, beta=0.440686793509772)
    m.init_(s, "random", rng=rng)
    n = 60
    m.sweep_(s, alg, n)
    best = 0.0
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        l0 = ctx.launch_count()
        check(lib().mcx_sweep(s.h_lat, n))
        e1.record(stream); torch.cuda.synchronize()
        best = max(best, n * L * L / (e0.elapsed_time(e1) * 1e6))
    print(json.dumps({"env": env, "attempts_per_ns": round(best, 1), "launches_per_sweep": (ctx.launch_count() - l0) / n}), flush=True)
    del s
