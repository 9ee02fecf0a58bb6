"""Blume-Capel 2-D sweep rate: vectorised k_bc2d vs the rows-of-8 kernel (MCX_BC2D=1 / 0)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcx_b200 as m
from mcx_b200._lib import check, lib

stream = torch.cuda.Stream()
ctx = m.Context(0, stream=stream.cuda_stream)
RULE = os.environ.get("BC_RULE", "metropolis")       # BC_RULE=heatbath: the two-threshold rule (blume_capel.jl:61-85)
for L, n in ((8192, 1), (1024, 64), (512, 256)):
    row = {"L": L, "chains": n, "rule": RULE}
    for mode in ("1", "0"):
        os.environ["MCX_BC2D"] = mode
        for track in (False, True):
            s = m.BlumeCapel([L, L], J=1, D=0.5, nchains=n, ctx=ctx); s.set_tracking(track)
            rng = m.PhiloxRNG(3); alg = (m.HeatBath if RULE == 'heatbath' else m.Metropolis)(rng, beta=0.9)
            m.init_(s, "random", rng=rng)
            ns = 20
            m.sweep_(s, alg, ns)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(stream)
            check(lib().mcx_sweep(s.h_lat, ns))
            e1.record(stream); torch.cuda.synchronize()
            row[("bc2d" if mode == "1" else "rows8") + ("_tracked" if track else "")] = round(ns * n * L * L / (e0.elapsed_time(e1) * 1e6), 1)
            del s
    print(json.dumps(row), flush=True)
