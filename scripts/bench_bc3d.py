"""3-D sweep rates: Blume-Capel (vectorised k_bc3d vs rows-of-8 kernel, MCX_BC2D=1 / 0) for the three rules."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcx_b200 as m
from mcx_b200._lib import check, lib

stream = torch.cuda.Stream()
ctx = m.Context(0, stream=stream.cuda_stream)
for L, n in ((256, 1), (128, 8), (64, 64)):
    for rule in ("metropolis", "heatbath"):
        row = {"dims": [L, L, L], "chains": n, "rule": rule}
        for mode in ("1", "0"):
            os.environ["MCX_BC2D"] = mode
            s = m.BlumeCapel([L, L, L], J=1, D=0.5, nchains=n, ctx=ctx); s.set_tracking(False)
            rng = m.PhiloxRNG(3); alg = (m.HeatBath if rule == "heatbath" else m.Metropolis)(rng, beta=0.5)
            m.init_(s, "random", rng=rng)
            ns = 10
            m.sweep_(s, alg, ns)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(stream)
            check(lib().mcx_sweep(s.h_lat, ns))
            e1.record(stream); torch.cuda.synchronize()
            row["bc3d" if mode == "1" else "rows8"] = round(ns * n * L ** 3 / (e0.elapsed_time(e1) * 1e6), 1)
            del s
        print(json.dumps(row), flush=True)
